/*
 * oracle/spirv_cpu.h — TEST INFRASTRUCTURE ONLY (parity oracle; never on the product path).
 *
 * CPU restatement of the shader stage of visor: the SPIR-V front end + JIT wrappers of
 * spirv_compile.cpp (reference file, 2452 lines).  The reference lowers SPIR-V to LLVM 6.0.0 IR
 * and JITs it; LLVM 6.0.0 is an un-vendored dependency (visor.vcxproj:116,122-123) that is absent
 * here, so this stage cannot be compiled from the reference.  It is restated as an interpreter that
 * follows spirv_compile.cpp pass by pass and opcode by opcode, including its bugs (see the
 * per-function citations in spirv_cpu.cpp).  The arithmetic the JIT emits is plain IEEE-754 binary32
 * with no fast-math flags (IRBuilder at spirv_compile.cpp:676 sets none), so evaluating the same
 * operations in the same order in C++ compiled with -ffp-contract=off is bit-identical for
 * + - * / sqrt; sin/cos/pow go to the MSVC CRT in the reference and are NOT reproducible.
 *
 * PARITY STATUS of this file: "parity unpinned" against the reference's own tests (it has none,
 * SURVEY.md §4/§8c) and against the real LLVM-JIT output (not buildable here).  It is pinned only
 * by hand-computed known answers and numpy float32 re-computations in tests/test_oracle_shader.py.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>

namespace vor
{
// Resource accessors the JIT'd wrappers call by name in the reference
// (GetVertexAttributeData / GetDescriptorBufferPointer / GetDescriptorImage /
//  GetPushConstantPointer / sample_tex_wrapped / sample_cube_wrapped — spirv_compile.cpp:552-627,
//  gpu.h:64-67).  `user` stands in for `const GPUState &state`.
struct ShaderEnv
{
  void *user;
  void (*vertex_attr)(void *user, uint32_t vertexIndex, uint32_t attr, float out[4]);
  const uint8_t *(*buffer_ptr)(void *user, uint32_t set, uint32_t bind);
  const void *(*image)(void *user, uint32_t set, uint32_t bind);
  const uint8_t *(*push_ptr)(void *user, uint32_t offset);
  void (*sample_tex)(void *user, float u, float v, const void *img, uint64_t byteOffs, float out[4]);
  void (*sample_cube)(void *user, float x, float y, float z, const void *img, float out[4]);
};

struct Module;
struct Entry;

// VertexCacheEntry (gpu.h:53-57) as 44 floats: position[4] then interps[10][4].
enum { kVertexFloats = 44 };

// CompileFunction (spirv_compile.cpp:645). NULL + *err on anything the reference would assert on.
Module *compile(const uint32_t *code, size_t words, std::string *err);
// accept the extended opcode set (see spirv_cpu.cpp) in subsequent compile() calls; off by default
void set_extended(bool on);
// GetFuncPointer (spirv_compile.cpp:2434): entry by OpEntryPoint name.
const Entry *find_entry(const Module *m, const char *name);
void destroy(Module *m);
int entry_stage(const Entry *e);    // 0 vertex, 4 fragment

// The exported VS wrapper (spirv_compile.cpp:1912-2117): void vs(state, vertexIndex, VertexCacheEntry&).
// Slots the shader does not write are left untouched in `out`.
void run_vertex(const Entry *e, const ShaderEnv &env, uint32_t vertexIndex, float out[kVertexFloats]);
// The exported FS wrapper (spirv_compile.cpp:2118-2366): void fs(state, pixdepth, bary, tri[3], out).
// Returns true when the invocation executed OpKill (extended mode only; `out` is then left untouched).
bool run_fragment(const Entry *e, const ShaderEnv &env, float pixdepth, const float bary[4],
                  const float *tri /* 3 x kVertexFloats */, float out[4]);

// GetVertexAttributeData's format switch (spirv_compile.cpp:576-626) on an already-resolved pointer.
// Returns false for formats the reference asserts on.
bool fetch_vertex_attr(uint32_t vkFormat, const uint8_t *ptr, float out[4]);
}    // namespace vor
