// spirv_ptx.h — SPIR-V -> PTX lowering for the B200 path.
//
// Replaces the LLVM-6 JIT front/back end of the reference (spirv_compile.cpp:645-2432): the same
// SPIR-V subset is lowered to one `.visible .func` PTX device function per
// entry point — `vb200_vs` for vertex, `vb200_fs` for fragment entry points — which the runtime splices
// into the PTX of the hand-written kernel that calls it (scaffold.cu) and compiles as one module.  Every float operation is emitted with an
// explicit `.rn` rounding modifier, which ptxas never contracts into FMA, and without `.ftz`, so the
// arithmetic is IEEE-754 binary32 in the reference's operation order.
// Host-only code: no CUDA dependency, so it is unit-testable without a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace vb200
{
enum Stage
{
  STAGE_VERTEX = 0,
  STAGE_FRAGMENT = 4
};

struct ResourceSlot
{
  uint32_t set, binding;
  bool is_image;    // true: Vb200Env::images[slot]; false: Vb200Env::res[slot]
  uint32_t slot;
};

struct ShaderEntry
{
  std::string name;
  int stage = 0;
  std::string ptx;                        // complete PTX module defining vb200_vs or vb200_fs
  std::vector<ResourceSlot> resources;    // descriptors this entry dereferences
  uint32_t attr_mask = 0;                 // VS: vertex attribute locations fetched
  uint32_t out_slot_mask = 0;             // VS: interpolant slots written
  uint32_t in_slot_mask = 0;              // FS: interpolant slots read
  bool uses_push = false;
  bool uses_kill = false;                 // FS (extended mode): contains OpKill -> only the ordered tile kernel is exact
};

struct ShaderModule
{
  std::vector<ShaderEntry> entries;
};

// CompileFunction (spirv_compile.cpp:645). Returns NULL and fills *err on anything outside the
// reference's subset (where the reference asserts).
ShaderModule *compile_spirv(const uint32_t *code, size_t words, std::string *err);
// option "extended_spirv": accept a few opcodes beyond the reference's subset in later compile_spirv calls
void set_extended_spirv(bool on);

// VS descriptors live in Vb200Env::res[0..7], FS descriptors in res[8..15], FS images in images[0..7].
enum
{
  kVsResBase = 0,
  kFsResBase = 8,
  kResPerStage = 8
};
}    // namespace vb200
