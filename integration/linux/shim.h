/*
 * integration/linux/shim.h — force-included (-include) before every UNMODIFIED reference translation unit
 * so the MSVC-flavoured sources compile with g++ on Linux.  It only supplies
 * headers the reference gets implicitly from <windows.h>/MSVC and neutralises __declspec. Used by the
 * Linux build of the CUDA ICD (integration/Makefile) and by the checker's build of the reference
 * (oracle/Makefile).
 */
#pragma once
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include <float.h>
#include <math.h>
#include <atomic>
#include <mutex>
#define __declspec(x) __DECLSPEC_##x
#define __DECLSPEC_dllexport
#define __DECLSPEC_thread thread_local
#define __DECLSPEC_align(n) alignas(n)
#define sprintf_s(buf, ...) snprintf(buf, sizeof(buf), __VA_ARGS__)
enum BC4Mode : int;
enum BC5Mode : int;
