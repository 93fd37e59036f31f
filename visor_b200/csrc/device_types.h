// device_types.h — structures shared by the host runtime, the CUDA kernels and the PTX the shader
// compiler emits.  Offsets used by generated PTX are spelled out as VB200_ENV_* constants and
// static_assert'ed against the struct so the three can never drift apart.
#pragma once
#include <stddef.h>
#include <stdint.h>

#define VB200_TILE 32            // screen tile edge in pixels (the reference's blockSize, rasterizer.cpp:454)
// Packed tile range of a triangle: tx0 | ty0 << 8 | tx1 << 16 | ty1 << 24 (inclusive). A live range has
// tx0 <= tx1, so a value with tx0 > tx1 can never be one: it marks culled / off-screen triangles. (All
// ones would collide with a triangle that lies entirely in tile (255, 255) of an 8192 x 8192 target.)
#define VB200_TILES_DEAD 0x000000ffu
#define VB200_MAX_SLOTS 10       // interpolant float4 slots per vertex (VertexCacheEntry::interps, gpu.h:56)
#define VB200_MAX_RES 16         // descriptor slots a pipeline may reference
#define VB200_MAX_IMAGES 8

struct Vb200Attr    // VkPipeline_T::vattrs (precompiled.h:118-124)
{
  uint32_t format, stride, offset, vb;
};

struct Vb200Image    // device-side view of VkImage_T (precompiled.h:89-98)
{
  const uint8_t *pixels;    // device address of VkImage_T::pixels
  uint32_t width, height;
  uint32_t bpp, format;
  uint32_t mips, layers;
  uint64_t slice_bytes;    // full mip-chain size of one layer (CalcSubresourceByteOffset, precompiled.cpp:18-33)
};

// Everything a shader invocation can reach: the device-side GPUState (gpu.h:3-23).
// Passed to the kernels as a __grid_constant__ parameter; the PTX shader functions get its address.
struct alignas(16) Vb200Env    // 16-byte aligned: generated PTX reads push constants with v4 loads
{
  const uint8_t *vb[4];               // vbs[i].buffer->bytes + vbs[i].offset, as device addresses
  Vb200Attr attrs[16];
  const uint8_t *res[VB200_MAX_RES];  // UBO slot -> bufferInfo.buffer->bytes + offset (device address)
  uint8_t push[128];                  // GPUState::pushconsts
  Vb200Image images[VB200_MAX_IMAGES];
};

#define VB200_ENV_VB 0
#define VB200_ENV_ATTRS 32
#define VB200_ENV_RES 288
#define VB200_ENV_PUSH 416
#define VB200_ENV_IMAGES 544
#define VB200_IMAGE_SIZE 40

#ifdef __cplusplus
static_assert(offsetof(Vb200Env, vb) == VB200_ENV_VB, "env layout");
static_assert(offsetof(Vb200Env, attrs) == VB200_ENV_ATTRS, "env layout");
static_assert(offsetof(Vb200Env, res) == VB200_ENV_RES, "env layout");
static_assert(offsetof(Vb200Env, push) == VB200_ENV_PUSH, "env layout");
static_assert(offsetof(Vb200Env, images) == VB200_ENV_IMAGES, "env layout");
static_assert(sizeof(Vb200Image) == VB200_IMAGE_SIZE, "image layout");
static_assert(sizeof(Vb200Env) <= 1024, "env must stay a cheap kernel parameter");
#endif

// Per-vertex raster record written by the vertex kernel: ToWindow (rasterizer.cpp:248-249) and the
// per-vertex part of triangle setup (invw, depth = z*invw, rasterizer.cpp:449-452).
struct Vb200RasterVertex
{
  int32_t x, y;
  float invw, depth;
};

// Per-triangle record (16 B) written by the setup kernel, read by the tile kernels: the triangle's three
// post-VS record slots and 1/|area2|. Everything else a tile kernel needs is a 16-byte gather from the
// L2-resident Vb200RasterVertex array per corner, so a triangle costs 16 B of HBM instead of 64.
struct Vb200TriRecord
{
  uint32_t s0, s1, s2;    // post-VS record slot of each corner
  float invarea;          // 1.0f / float(|area2|) (rasterizer.cpp:448); undefined for dead triangles
};

// The same triangle as the tile kernels hold it in registers (record + the three vertex gathers).
struct Vb200TriSetup
{
  int32_t x0, y0, x1, y1, x2, y2;    // window coordinates of the three corners
  float invw0, invw1, invw2;
  float d0, d1, d2;
  uint32_t s0, s1, s2;               // post-VS record slot of each corner
  float invarea;                     // 1.0f / float(|area2|) (rasterizer.cpp:448); undefined for dead triangles
};
#ifdef __cplusplus
static_assert(sizeof(Vb200TriRecord) == 16, "triangle record");
#endif

// Statistics counters. Same-address atomics serialise in the L2 atomic unit (~0.6 ns each), so every
// counter is spread over 32 slots (slot = blockIdx & 31, 128-byte stride) and summed on read.
#define VB200_COUNTER_SLOTS 32
struct Vb200DrawCounters
{
  struct alignas(128) Slot
  {
    unsigned long long triangles_out, fragments_covered, fragments_shaded, tile_pairs;
  } slot[VB200_COUNTER_SLOTS];
};

// Fixed-function state of one draw, uniform across the grid.
struct Vb200RasterState
{
  uint32_t width, height;
  uint32_t tiles_x, tiles_y;
  uint32_t tiles_x_magic;    // floor(2^32 / tiles_x) + 1: tile / tiles_x == __umulhi(tile, magic) for tile < 65536
  uint32_t depth_op;         // VkCompareOp; 7 (ALWAYS) or no depth image -> no test
  uint32_t depth_write;
  uint32_t has_depth;
  uint32_t blend_enable, src_factor, dst_factor, blend_op;
  uint32_t nslots;           // float4 slots per post-VS vertex record
  uint32_t owner_rank, owner_world;    // sort-first tile ownership (tile % world == rank)
  uint32_t count_fragments;
  uint32_t color_bpp;
  // resolve kernels: the visibility key's low word carries (triangle id << 8 | record slot), so the
  // shading pass finds the winner's record in shared memory. Needs triangle ids below 2^24.
  uint32_t slot_keys;
};
