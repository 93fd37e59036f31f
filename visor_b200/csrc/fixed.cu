// fixed.cu — the shader-independent sm_100a kernels of the draw path.
//
//   K0  clear                     ClearTarget                      rasterizer.cpp:312-361
//   K2  triangle assembly+setup   GetIndex, ShadeVerts assembly,   rasterizer.cpp:100-232, 385-452
//                                 DrawTriangles setup / cull / bbox
//   K3  binning                   the per-triangle 32x32 block      rasterizer.cpp:454-515
//                                 split + FIFO queue, restated as
//                                 per-screen-tile lists appended by
//                                 the setup kernel itself (+ per-tile
//                                 sort by triangle id for the in-order path)
//   K5' stand-alone sampler       sample_tex_wrapped/_cube_wrapped texture_sampling.cpp:139-250
//   K6  tile pack / unpack        (sort-first multi-GPU; no reference equivalent)
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "kernels.h"
#include "raster_common.cuh"

namespace vb200
{
namespace
{
constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// K0: clears. 16-byte stores, grid-stride; grid sized to a multiple of the SM count by the caller.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_clear_u32(uint32_t *dst, uint32_t value, size_t count)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = count / 4;
  uint4 *d4 = (uint4 *)dst;
  const uint4 v4 = make_uint4(value, value, value, value);
  for(size_t k = i; k < n4; k += stride)
    d4[k] = v4;
  for(size_t k = n4 * 4 + i; k < count; k += stride)
    dst[k] = value;
}

__global__ void __launch_bounds__(kThreads) k_clear_u8(uint8_t *dst, uint8_t value, size_t count)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for(size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
    dst[k] = value;
}

// ------------------------------------------------------------------------------------------------
// index range: min/max of the index values a draw references (bounds the unique-vertex VS launch)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_index(const void *ib, uint32_t index_type, uint32_t i)
{
  // GetIndex (rasterizer.cpp:100-119)
  if(index_type == 0u)
    return (uint32_t)__ldg((const uint16_t *)ib + i);
  return __ldg((const uint32_t *)ib + i);
}

__global__ void __launch_bounds__(kThreads) k_index_range(const void *ib, uint32_t index_type, uint32_t first,
                                                         uint32_t count, uint32_t *range)
{
  // persistent grid-stride reduction: one atomic pair per CTA (the L2 atomic unit serialises per
  // address, so per-warp atomics on two words would dominate the kernel)
  __shared__ uint32_t s_lo[kThreads / 32], s_hi[kThreads / 32];
  uint32_t lo = 0xffffffffu, hi = 0u;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if(index_type != 0u && ((((uintptr_t)ib) + 4ull * first) & 15) == 0)
  {
    // 16-byte vector loads of 4 indices
    const uint4 *v4 = (const uint4 *)((const uint32_t *)ib + first);
    const uint32_t n4 = count / 4u;
    for(uint32_t i = tid; i < n4; i += stride)
    {
      const uint4 v = __ldg(v4 + i);
      lo = min(min(lo, v.x), min(min(v.y, v.z), v.w));
      hi = max(max(hi, v.x), max(max(v.y, v.z), v.w));
    }
    for(uint32_t i = n4 * 4u + tid; i < count; i += stride)
    {
      const uint32_t v = load_index(ib, index_type, first + i);
      lo = min(lo, v);
      hi = max(hi, v);
    }
  }
  else
    for(uint32_t i = tid; i < count; i += stride)
    {
      const uint32_t v = load_index(ib, index_type, first + i);
      lo = min(lo, v);
      hi = max(hi, v);
    }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if((threadIdx.x & 31) == 0)
  {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if(threadIdx.x < 32)
  {
    lo = threadIdx.x < kThreads / 32 ? s_lo[threadIdx.x] : 0xffffffffu;
    hi = threadIdx.x < kThreads / 32 ? s_hi[threadIdx.x] : 0u;
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if(threadIdx.x == 0 && lo <= hi)
    {
      atomicMin(&range[0], lo);
      atomicMax(&range[1], hi);
    }
  }
}

__global__ void k_init_range(uint32_t *range, uint32_t lo, uint32_t hi)
{
  range[0] = lo;
  range[1] = hi;
}

// ------------------------------------------------------------------------------------------------
// K2 + K3: triangle assembly, setup, cull, bbox AND binning, one pass over the triangles.
//
// Binning appends a surviving triangle's id straight to the list of every 32x32 tile its bbox touches
// (the reference's per-triangle block split + FIFO push, rasterizer.cpp:454-485). Every tile owns a
// fixed window of list_cap entries; the per-tile counter doubles as the append cursor and counts every
// append, so a tile that received more than list_cap triangles is recognisable afterwards and its tile
// kernel CTA falls back to scanning the packed tile ranges (tri_tiles) — exact, no host involvement.
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel of the frame's chain (vertex -> setup -> [sort] -> tiles) lets its
// successor be scheduled early (vb200_pdl_trigger when it starts) and waits for its predecessor's results before it
// touches them (vb200_pdl_wait; a no-op when the kernel was launched in plain stream order).
__device__ __forceinline__ void vb200_pdl_trigger()
{
  asm volatile("griddepcontrol.launch_dependents;");
}
__device__ __forceinline__ void vb200_pdl_wait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static void launch_dependent(void (*kernel)(KArgs...), uint32_t grid, cudaStream_t s, Args... args)
{
  static bool pdl = getenv("VB200_NO_PDL") == nullptr;
  if(!pdl)
  {
    kernel<<<grid, kThreads, 0, s>>>(args...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = s;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  if(e == cudaErrorNotSupported || e == cudaErrorInvalidValue)
  {
    // an environment without programmatic launches: plain stream order from now on
    cudaGetLastError();
    pdl = false;
    kernel<<<grid, kThreads, 0, s>>>(args...);
  }
}

// Tile ownership (sort-first): tile t belongs to rank t % world, and is the (t / world)-th tile of its owner.
// kOwner: 0 = a single GPU (no test at all), 1 = world is a power of two (mask and shift), 2 = any world (the
// integer division costs ~20 instructions per test; it used to be paid four times per triangle on one GPU too).
template <int kOwner>
__device__ __forceinline__ bool tile_owned(uint32_t tile, uint32_t rank, uint32_t world)
{
  if(kOwner == 0)
    return true;
  if(kOwner == 1)
    return (tile & (world - 1u)) == rank;
  return (tile % world) == rank;
}
template <int kOwner>
__device__ __forceinline__ uint32_t tile_slot(uint32_t tile, uint32_t world, uint32_t world_shift)
{
  if(kOwner == 0)
    return tile;
  if(kOwner == 1)
    return tile >> world_shift;
  return tile / world;
}

// Each thread sets up kSetupPerThread triangles, phase by phase, so that the index loads of all of them,
// then the vertex gathers of all of them, then the returning cursor atomics of all of them are in flight
// together (the kernel is a chain of three dependent memory operations per triangle and little else).
// kTables: the batch holds several draws (per-triangle lookup of the draw and its vertex span); a single draw
// compiles without that code — left in, predicated off, it was a sixth of the instructions the kernel issued.
template <int kSetupPerThread, bool kTables, int kOwner>
__global__ void __launch_bounds__(kThreads) k_setup(const Vb200SetupParams p)
{
  const uint32_t worldShift = kOwner == 1 ? (uint32_t)(__ffs((int)p.owner_world) - 1) : 0u;
  vb200_pdl_trigger();
  uint32_t t[kSetupPerThread], s0[kSetupPerThread], s1[kSetupPerThread], s2[kSetupPerThread];
  uint32_t tiles[kSetupPerThread];
  bool alive[kSetupPerThread];
  float invarea[kSetupPerThread];
  int2 va[kSetupPerThread], vb[kSetupPerThread], vc[kSetupPerThread];
  // ranges above 2x2 tiles are queued here and walked by the whole CTA (a full-screen triangle covers
  // thousands of tiles: one thread appending to all of them would take tens of microseconds)
  __shared__ uint32_t s_big[kThreads * kSetupPerThread][2];
  __shared__ uint32_t s_nbig;
  if(threadIdx.x == 0)
    s_nbig = 0;
  __syncthreads();

  // ---- 1. triangle assembly (rasterizer.cpp:128-232): list = (3t, 3t+1, 3t+2); strip alternates
  // (t, t+1, t+2) / (t+1, t, t+2) to preserve winding; GetIndex (:100-119). In a batch of several draws the
  // triangle's draw is the last one whose tri_base is <= t (a handful of L1-resident probes).
  Vb200VertexSpan span[kSetupPerThread];
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
  {
    t[k] = (blockIdx.x * kSetupPerThread + k) * blockDim.x + threadIdx.x;
    alive[k] = t[k] < p.num_tris;
    tiles[k] = VB200_TILES_DEAD;
    invarea[k] = 0.0f;
    s0[k] = s1[k] = s2[k] = 0u;
    span[k] = p.span0;
    if(alive[k])
    {
      Vb200BatchDraw d = p.draw0;
      if(kTables)
      {
        uint32_t lo = 0, hi = p.num_draws;    // invariant: draws[lo].tri_base <= t < draws[hi].tri_base
        while(hi - lo > 1u)
        {
          const uint32_t mid = (lo + hi) >> 1;
          if(__ldg(&p.draws[mid].tri_base) <= t[k])
            lo = mid;
          else
            hi = mid;
        }
        d = p.draws[lo];
        span[k] = p.spans[d.span];
      }
      const uint32_t local = t[k] - d.tri_base;
      uint32_t c0, c1, c2;
      if(p.topology == 3u)
      {
        c0 = d.first + 3u * local;
        c1 = c0 + 1u;
        c2 = c0 + 2u;
      }
      else
      {
        const uint32_t b = d.first + local;
        c0 = (local & 1u) ? b + 1u : b;
        c1 = (local & 1u) ? b : b + 1u;
        c2 = b + 2u;
      }
      if(d.ib)
      {
        c0 = load_index(d.ib, d.index_type, c0);
        c1 = load_index(d.ib, d.index_type, c1);
        c2 = load_index(d.ib, d.index_type, c2);
        // a malformed index beyond the bound vertex buffers kills the triangle
        alive[k] = c0 < p.vertex_bound && c1 < p.vertex_bound && c2 < p.vertex_bound;
      }
      s0[k] = c0;
      s1[k] = c1;
      s2[k] = c2;
    }
  }
  if(p.range)    // the lone draw whose span was measured on the device: [min, max], at most span0.count records
  {
    const uint32_t lo = p.range[0], hi = p.range[1];
#pragma unroll
    for(int k = 0; k < kSetupPerThread; k++)
    {
      span[k].src_base = lo;
      span[k].count = hi >= lo ? min(hi - lo + 1u, span[k].count) : 0u;
    }
  }
  // Everything above read inputs only (draw table, index buffer, a device-measured range from before the vertex
  // kernel): with a programmatic launch it ran while the vertex kernel's last wave drained. From here on the
  // vertex kernel's raster records and zeroed tile counters are needed.
  vb200_pdl_wait();
  // ---- 2. window positions of the three corners (the raster record's first 8 bytes)
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
  {
    s0[k] -= span[k].src_base;
    s1[k] -= span[k].src_base;
    s2[k] -= span[k].src_base;
    alive[k] = alive[k] && s0[k] < span[k].count && s1[k] < span[k].count && s2[k] < span[k].count;
    s0[k] += span[k].slot_base;
    s1[k] += span[k].slot_base;
    s2[k] += span[k].slot_base;
    va[k] = vb[k] = vc[k] = make_int2(0, 0);
    if(alive[k])
    {
      va[k] = __ldg((const int2 *)(p.rv + s0[k]));
      vb[k] = __ldg((const int2 *)(p.rv + s1[k]));
      vc[k] = __ldg((const int2 *)(p.rv + s2[k]));
    }
  }
  // ---- 3. double_triarea (rasterizer.cpp:272-275), zero-area skip (:398), facing / cull (:401-424),
  // MinMax + clamp (:428-435; the pixel loops run over [min, max), :538-540)
  uint32_t survivors = 0;
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
  {
    if(alive[k])
    {
      const int area2 = (vb[k].x - va[k].x) * (vc[k].y - va[k].y) - (vb[k].y - va[k].y) * (vc[k].x - va[k].x);
      // 1.0f / float(|area2|) (rasterizer.cpp:448): the dedicated reciprocal returns the same bits as the general
      // division (both are the correctly rounded quotient) in a shorter sequence
      invarea[k] = __frcp_rn((float)(area2 < 0 ? -area2 : area2));
      const int flipped = (p.front_face == 1u) ? -area2 : area2;
      if(area2 == 0)
        alive[k] = false;
      else if(flipped > 0 && (p.cull_mode & 1u))
        alive[k] = false;
      else if(flipped < 0 && (p.cull_mode & 2u))
        alive[k] = false;
    }
    survivors += alive[k] ? 1u : 0u;
    if(alive[k])
    {
      const int minx = max(0, min(va[k].x, min(vb[k].x, vc[k].x)));
      const int miny = max(0, min(va[k].y, min(vb[k].y, vc[k].y)));
      const int maxx = min((int)p.width - 1, max(va[k].x, max(vb[k].x, vc[k].x)));
      const int maxy = min((int)p.height - 1, max(va[k].y, max(vb[k].y, vc[k].y)));
      if(minx < maxx && miny < maxy)
        tiles[k] = (uint32_t)(minx / VB200_TILE) | ((uint32_t)(miny / VB200_TILE) << 8) |
                   ((uint32_t)((maxx - 1) / VB200_TILE) << 16) | ((uint32_t)((maxy - 1) / VB200_TILE) << 24);
      else
        alive[k] = false;
    }
  }

  // ---- 4. binning. Neighbouring triangles of a mesh land in the same tile, so per-lane atomics on the
  // tile cursors would serialise in the L2 atomic unit: __match_any_sync groups the lanes of a warp that
  // target the same tile and one lane per group reserves room for all of them. Ranges of up to 2x2 tiles
  // (every triangle smaller than a tile) are handled quadrant by quadrant; the cursor atomics of both of a
  // thread's triangles are issued before either result is used.
  const uint32_t lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
  uint32_t nx[kSetupPerThread], ny[kSetupPerThread];
  bool used[kSetupPerThread];    // some tile of this rank needs the triangle
  uint32_t mypairs = 0;
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
  {
    nx[k] = alive[k] ? ((tiles[k] >> 16) & 0xffu) - (tiles[k] & 0xffu) + 1u : 0u;
    ny[k] = alive[k] ? (tiles[k] >> 24) - ((tiles[k] >> 8) & 0xffu) + 1u : 0u;
    used[k] = false;
  }
#pragma unroll
  for(uint32_t q = 0; q < 4u; q++)
  {
    const uint32_t qx = q & 1u, qy = q >> 1;
    uint32_t tile[kSetupPerThread], peers[kSetupPerThread], pos[kSetupPerThread];
    bool any = false;
#pragma unroll
    for(int k = 0; k < kSetupPerThread; k++)
    {
      tile[k] = 0xffffffffu;
      if(nx[k] <= 2u && ny[k] <= 2u && qx < nx[k] && qy < ny[k])
      {
        tile[k] = (((tiles[k] >> 8) & 0xffu) + qy) * p.tiles_x + (tiles[k] & 0xffu) + qx;
        if(!tile_owned<kOwner>(tile[k], p.owner_rank, p.owner_world))
          tile[k] = 0xffffffffu;
      }
      any |= tile[k] != 0xffffffffu;
    }
    if(!__any_sync(0xffffffffu, any))
      continue;    // most triangles touch one tile: the other three quadrants are empty for the whole warp
#pragma unroll
    for(int k = 0; k < kSetupPerThread; k++)
    {
      peers[k] = __match_any_sync(0xffffffffu, tile[k]);
      pos[k] = 0;
      if(tile[k] != 0xffffffffu && (uint32_t)(__ffs(peers[k]) - 1) == lane)
        pos[k] = atomicAdd(&p.tile_count[tile[k]], (uint32_t)__popc(peers[k]));
    }
#pragma unroll
    for(int k = 0; k < kSetupPerThread; k++)
    {
      pos[k] = __shfl_sync(0xffffffffu, pos[k], __ffs(peers[k]) - 1) + __popc(peers[k] & below);
      if(tile[k] != 0xffffffffu)
      {
        used[k] = true;
        mypairs++;
        if(pos[k] < p.list_cap)
          p.list[(size_t)tile_slot<kOwner>(tile[k], p.owner_world, worldShift) * p.list_cap + pos[k]] = t[k];
      }
    }
  }
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
    if(alive[k] && (nx[k] > 2u || ny[k] > 2u))
    {
      used[k] = true;
      const uint32_t slot = atomicAdd(&s_nbig, 1u);
      s_big[slot][0] = tiles[k];
      s_big[slot][1] = t[k];
    }
  // ---- 5. the 16-byte triangle record (only of triangles this rank will rasterise: with sort-first
  // ownership that is 1/world of them) and the packed tile range (of all: the tile kernels' fallback list)
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
    if(t[k] < p.num_tris)
    {
      if(used[k])
        *(int4 *)(p.tri + t[k]) = make_int4((int)s0[k], (int)s1[k], (int)s2[k], __float_as_int(invarea[k]));
      p.tri_tiles[t[k]] = alive[k] ? tiles[k] : VB200_TILES_DEAD;
    }
  __syncthreads();
  const uint32_t nbig = s_nbig;
  for(uint32_t b = 0; b < nbig; b++)
  {
    const uint32_t bt = s_big[b][0], b_tri = s_big[b][1];
    const uint32_t b_tx0 = bt & 0xffu, b_ty0 = (bt >> 8) & 0xffu;
    const uint32_t b_nx = ((bt >> 16) & 0xffu) - b_tx0 + 1u, b_nt = b_nx * ((bt >> 24) - b_ty0 + 1u);
    for(uint32_t k = threadIdx.x; k < b_nt; k += kThreads)
    {
      const uint32_t tile = (b_ty0 + k / b_nx) * p.tiles_x + b_tx0 + k % b_nx;
      if(tile_owned<kOwner>(tile, p.owner_rank, p.owner_world))
      {
        const uint32_t pos = atomicAdd(&p.tile_count[tile], 1u);
        mypairs++;
        if(pos < p.list_cap)
          p.list[(size_t)tile_slot<kOwner>(tile, p.owner_world, worldShift) * p.list_cap + pos] = b_tri;
      }
    }
  }
  // statistics: one atomic pair per warp on counters spread over 32 slots (per-warp atomics on ONE word cost
  // ~20 us per million triangles: same-address atomics serialise in L2)
  mypairs = __reduce_add_sync(0xffffffffu, mypairs);
  survivors = __reduce_add_sync(0xffffffffu, survivors);
  if(lane == 0)
  {
    Vb200DrawCounters::Slot &c = p.counters->slot[(blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) & (VB200_COUNTER_SLOTS - 1)];
    if(survivors)
      atomicAdd(&c.triangles_out, (unsigned long long)survivors);
    if(mypairs)
      atomicAdd(&c.tile_pairs, (unsigned long long)mypairs);
  }
}

// Per-tile sort by triangle id: restores submission order after the unordered atomic append, which
// is what makes every pixel see its fragments in draw order (the serial FIFO of the reference).
// Padding-friendly bitonic network (all compare-exchanges put the minimum at the lower index, so
// virtual +inf elements past n never move).
__device__ __forceinline__ void cmpxchg(uint32_t *a, uint32_t lo, uint32_t hi)
{
  const uint32_t x = a[lo], y = a[hi];
  if(x > y)
  {
    a[lo] = y;
    a[hi] = x;
  }
}

// Force-inlined so that a call on the __shared__ staging array compiles to LDS/STS; strides are powers
// of two, so every index is shifts and masks.
__device__ __forceinline__ void bitonic_sort(uint32_t *a, uint32_t n)
{
  uint32_t logN = 0;
  while((1u << logN) < n)
    logN++;
  const uint32_t half = (1u << logN) >> 1;
  for(uint32_t lk = 1; lk <= logN; lk++)    // k = 1 << lk
  {
    const uint32_t k = 1u << lk, hk = k >> 1;
    for(uint32_t i = threadIdx.x; i < half; i += blockDim.x)
    {
      const uint32_t base = (i >> (lk - 1u)) << lk, pos = i & (hk - 1u);
      const uint32_t lo = base + pos, hi = base + (k - 1u - pos);
      if(hi < n)
        cmpxchg(a, lo, hi);
    }
    __syncthreads();
    for(uint32_t lj = lk - 1u; lj-- > 0u;)    // j = 1 << lj, from k/4 down to 1
    {
      const uint32_t j = 1u << lj;
      for(uint32_t i = threadIdx.x; i < half; i += blockDim.x)
      {
        const uint32_t lo = ((i >> lj) << (lj + 1u)) + (i & (j - 1u)), hi = lo + j;
        if(hi < n)
          cmpxchg(a, lo, hi);
      }
      __syncthreads();
    }
  }
}

constexpr uint32_t kSortSmem = 8192;    // entries (32 KB)

// Lists of up to 512 ids (two per thread, held in registers): stable LSD radix sort, four bits per pass,
// over as many bits as the largest id of the list has. Element i belongs to chunk i / 32, i.e. to one
// warp: MATCH.ANY on the digit gives its rank among the chunk's equal digits and, on the first lane of every
// group, the group's size; a scan over the 16 x 16 (digit-major, chunk-minor) counts turns that into
// positions. Padding (0xffffffff) has digit 15 in every pass and starts behind everything, so it stays there.
// The counts of one digit are 17 words apart, not 16: a warp's lanes address them by digit, and 16 would put
// all even digits into one bank. `s`: 512 keys, 16 x 17 counts, 8 warp totals, 1 word for the OR of all ids.
__device__ __forceinline__ void radix_sort_512(uint32_t *a, uint32_t n, uint32_t *s)
{
  uint32_t *buf = s, *hist = s + 512, *wtot = s + 784, *bitsOr = s + 792;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t below = (1u << lane) - 1u;
  const bool in0 = tid < n, in1 = tid + 256u < n;
  uint32_t k0 = in0 ? a[tid] : 0xffffffffu, k1 = in1 ? a[tid + 256u] : 0xffffffffu;
  if(tid == 0)
    *bitsOr = 0u;
  __syncthreads();
  const uint32_t mine = __reduce_or_sync(0xffffffffu, (in0 ? k0 : 0u) | (in1 ? k1 : 0u));
  if(lane == 0 && mine)
    atomicOr(bitsOr, mine);
  __syncthreads();
  const uint32_t bits = 32u - (uint32_t)__clz((int)*bitsOr);
  for(uint32_t shift = 0; shift < bits; shift += 4u)
  {
    const uint32_t d0 = (k0 >> shift) & 15u, d1 = (k1 >> shift) & 15u;
    const uint32_t own = (tid >> 4) * 17u + (tid & 15u);    // this thread's count in the scan
    hist[own] = 0u;
    __syncthreads();
    const uint32_t m0 = __match_any_sync(0xffffffffu, d0), m1 = __match_any_sync(0xffffffffu, d1);
    const uint32_t r0 = __popc(m0 & below), r1 = __popc(m1 & below);
    if(r0 == 0u)
      hist[d0 * 17u + warp] = __popc(m0);
    if(r1 == 0u)
      hist[d1 * 17u + warp + 8u] = __popc(m1);
    __syncthreads();
    // exclusive scan of the 256 counts, one per thread
    const uint32_t v = hist[own];
    uint32_t x = v;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if(lane >= (uint32_t)o)
        x += y;
    }
    if(lane == 31u)
      wtot[warp] = x;
    __syncthreads();
    const uint32_t base = __reduce_add_sync(0xffffffffu, lane < warp ? wtot[lane] : 0u);    // warp < 8
    hist[own] = base + x - v;
    __syncthreads();
    buf[hist[d0 * 17u + warp] + r0] = k0;
    buf[hist[d1 * 17u + warp + 8u] + r1] = k1;
    __syncthreads();
    k0 = buf[tid];
    k1 = buf[tid + 256u];
  }
  if(in0)
    a[tid] = k0;
  if(in1)
    a[tid + 256u] = k1;
}

__global__ void __launch_bounds__(kThreads) k_sort(uint32_t *list, const uint32_t *tile_count, uint32_t list_cap,
                                                  uint32_t rank, uint32_t world, uint32_t ntiles, uint32_t rankSort)
{
  __shared__ __align__(16) uint32_t s[kSortSmem];
  vb200_pdl_trigger();
  vb200_pdl_wait();    // the setup kernel's lists
  const uint32_t tile = blockIdx.x * world + rank;    // the grid holds the tiles this rank owns
  if(tile >= ntiles)
    return;
  const uint32_t n = tile_count[tile];
  if(n < 2u || n > list_cap)    // an overflowed list is not used: the tile kernel scans tri_tiles, in order
    return;
  uint32_t *a = list + (size_t)blockIdx.x * list_cap;
  int unsorted = 0;
  for(uint32_t i = threadIdx.x; i + 1u < n; i += blockDim.x)
    unsorted |= a[i] > a[i + 1u];
  if(!__syncthreads_or(unsorted))
    return;
  if(n <= 512u && !rankSort)
    radix_sort_512(a, n, s);
  else if(n <= 512u)
  {
    // (the earlier version, kept behind VB200_SORT_RANK=1 for comparison) rank sort. Triangle ids are
    // unique, so an id's final position is the number of smaller ids; every thread counts that for its
    // own elements against broadcast reads of the staged list — no barriers, no data-dependent branches.
    const uint32_t n4 = (n + 3u) & ~3u;
    for(uint32_t i = threadIdx.x; i < n4; i += blockDim.x)
      s[i] = i < n ? a[i] : 0xffffffffu;
    __syncthreads();
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
      const uint32_t e = s[i];
      uint32_t r = 0;
      for(uint32_t j = 0; j < n4; j += 4u)
      {
        const uint4 v = *(const uint4 *)(s + j);
        r += (v.x < e) + (v.y < e) + (v.z < e) + (v.w < e);
      }
      a[r] = e;
    }
  }
  else if(n <= kSortSmem)
  {
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      s[i] = a[i];
    __syncthreads();
    bitonic_sort(s, n);
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      a[i] = s[i];
  }
  else
    bitonic_sort(a, n);    // rare: > 8192 triangles over one tile; sorted in place through L2
}

// ------------------------------------------------------------------------------------------------
// stand-alone sampler (parity tests of the texture unit)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_sample(const Vb200Image img, int cube, unsigned long long byte_offset,
                                                    const float *uvw, float4 *out, size_t count)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= count)
    return;
  if(cube)
    out[i] = vb200_sample_cube_impl(uvw[3 * i], uvw[3 * i + 1], uvw[3 * i + 2], &img);
  else
    out[i] = vb200_sample_tex_impl(uvw[2 * i], uvw[2 * i + 1], &img, byte_offset);
}

// ------------------------------------------------------------------------------------------------
// K6: sort-first tile exchange. Tile t is owned by rank t % world; the k-th owned tile of a rank is
// t = rank + k*world and occupies 4096 bytes (32 rows x 128 B) at slot k of the rank's send buffer.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_tiles_pack(const uint32_t *color, uint32_t width, uint32_t height,
                                                        uint32_t tiles_x, uint32_t ntiles, uint32_t rank,
                                                        uint32_t world, uint32_t *dst)
{
  const uint32_t k = blockIdx.x;
  const uint32_t tile = rank + k * world;
  if(tile >= ntiles)
  {
    // the last slot of a rank that owns one tile fewer: defined contents for the all-gather
    for(uint32_t i = threadIdx.x; i < VB200_TILE * VB200_TILE; i += blockDim.x)
      dst[(size_t)k * 1024u + i] = 0u;
    return;
  }
  const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
  for(uint32_t i = threadIdx.x; i < VB200_TILE * VB200_TILE; i += blockDim.x)
  {
    const uint32_t x = tx * VB200_TILE + (i & 31u), y = ty * VB200_TILE + (i >> 5);
    dst[(size_t)k * 1024u + i] = (x < width && y < height) ? color[(size_t)y * width + x] : 0u;
  }
}

__global__ void __launch_bounds__(kThreads) k_tiles_unpack(uint32_t *color, uint32_t width, uint32_t height,
                                                          uint32_t tiles_x, uint32_t ntiles, uint32_t world,
                                                          uint32_t slots_per_rank, const uint32_t *src)
{
  const uint32_t tile = blockIdx.x;
  if(tile >= ntiles)
    return;
  const uint32_t owner = tile % world, k = tile / world;
  const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
  const uint32_t *s = src + ((size_t)owner * slots_per_rank + k) * 1024u;
  for(uint32_t i = threadIdx.x; i < VB200_TILE * VB200_TILE; i += blockDim.x)
  {
    const uint32_t x = tx * VB200_TILE + (i & 31u), y = ty * VB200_TILE + (i >> 5);
    if(x < width && y < height)
      color[(size_t)y * width + x] = s[i];
  }
}

// ------------------------------------------------------------------------------------------------
// K7: sort-first plumbing between the GPUs of one node (no reference equivalent)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_mgpu_barrier(const Vb200PeerSet flags, uint32_t rank, uint32_t world,
                                                    uint32_t epoch, uint32_t *timed_out)
{
  const uint32_t p = threadIdx.x;
  if(p >= world)
    return;
  // everything this GPU wrote before (earlier kernels of the stream, including their stores into peer and
  // multicast mappings) is ordered before the flag
  __threadfence_system();
  volatile uint32_t *theirs = (volatile uint32_t *)flags.peer[p] + rank;
  *theirs = epoch;
  __threadfence_system();
  volatile uint32_t *mine = (volatile uint32_t *)flags.peer[rank] + p;
  const long long t0 = clock64();
  while((int32_t)(*mine - epoch) < 0)
  {
    if(clock64() - t0 > 4000000000ll)    // ~2 s at 2 GHz
    {
      *timed_out = 1u;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(kThreads) k_mgpu_push(const Vb200PeerSet bufs, uint4 *multicast, uint32_t rank,
                                                       uint32_t world, size_t first16, size_t count16)
{
  const uint4 *src = (const uint4 *)bufs.peer[rank] + first16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count16; i += stride)
  {
    const uint4 v = src[i];
    if(multicast)
      asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(multicast + first16 + i),
                   "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                   "f"(__uint_as_float(v.w))
                   : "memory");
    else
      for(uint32_t r = 0; r < world; r++)
        if(r != rank)
          ((uint4 *)bufs.peer[r])[first16 + i] = v;
  }
}

int sm_count()
{
  static int n = 0;
  if(!n)
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if(n <= 0)
      n = 148;
  }
  return n;
}

uint32_t grid_for(size_t work_items, int per_thread = 1)
{
  // grid-stride kernels: a whole number of waves of (SM count x 8 CTAs), capped by the work
  const size_t need = (work_items + (size_t)kThreads * per_thread - 1) / ((size_t)kThreads * per_thread);
  const size_t wave = (size_t)sm_count() * 8;
  const size_t g = need < wave * 4 ? need : wave * 4;
  return (uint32_t)(g ? g : 1);
}
}    // namespace

int launch_clear_u32(uint32_t *dst, uint32_t value, size_t count, cudaStream_t s)
{
  if(!count)
    return 0;
  k_clear_u32<<<grid_for(count, 16), kThreads, 0, s>>>(dst, value, count);
  return 1;
}

int launch_clear_u8(uint8_t *dst, uint8_t value, size_t count, cudaStream_t s)
{
  if(!count)
    return 0;
  k_clear_u8<<<grid_for(count, 16), kThreads, 0, s>>>(dst, value, count);
  return 1;
}

int launch_index_range(const void *ib, uint32_t index_type, uint32_t first, uint32_t count, uint32_t *range,
                       cudaStream_t s)
{
  k_init_range<<<1, 1, 0, s>>>(range, 0xffffffffu, 0u);
  if(!count)
    return 1;
  // one wave: 4 CTAs per SM
  const uint32_t g = (uint32_t)std::min<size_t>((size_t)sm_count() * 4, (count + kThreads - 1) / kThreads);
  k_index_range<<<g ? g : 1, kThreads, 0, s>>>(ib, index_type, first, count, range);
  return 2;
}

int launch_setup(const Vb200SetupParams &p, cudaStream_t s)
{
  if(!p.num_tris)
    return 0;
  // triangles per thread: 2 (default) or 4 (VB200_SETUP_PER_THREAD, tuning aid)
  static const int perThread = []() {
    const char *e = getenv("VB200_SETUP_PER_THREAD");
    return e && atoi(e) == 4 ? 4 : 2;
  }();
  const uint32_t per_cta = kThreads * (uint32_t)perThread;
  const uint32_t grid = (p.num_tris + per_cta - 1) / per_cta;
  const bool tables = p.draws != nullptr;
  const int owner = p.owner_world <= 1u ? 0 : ((p.owner_world & (p.owner_world - 1u)) == 0u ? 1 : 2);
#define VB200_SETUP_CASE(N, T, O)                       \
  if(perThread == N && tables == T && owner == O)       \
    launch_dependent(k_setup<N, T, O>, grid, s, p);
#define VB200_SETUP_CASES(N) \
  VB200_SETUP_CASE(N, false, 0) VB200_SETUP_CASE(N, false, 1) VB200_SETUP_CASE(N, false, 2) \
  VB200_SETUP_CASE(N, true, 0) VB200_SETUP_CASE(N, true, 1) VB200_SETUP_CASE(N, true, 2)
  VB200_SETUP_CASES(2)
  VB200_SETUP_CASES(4)
#undef VB200_SETUP_CASES
#undef VB200_SETUP_CASE
  return 1;
}

int launch_sort(uint32_t *list, const uint32_t *tile_count, uint32_t list_cap, uint32_t rank, uint32_t world,
                uint32_t ntiles, cudaStream_t s)
{
  static const uint32_t rankSort = getenv("VB200_SORT_RANK") != nullptr;
  launch_dependent(k_sort, (ntiles + world - 1) / world, s, list, tile_count, list_cap, rank, world, ntiles, rankSort);
  return 1;
}

int launch_sample(const Vb200Image &img, int cube, uint64_t byte_offset, const float *uvw, float4 *out,
                  size_t count, cudaStream_t s)
{
  if(!count)
    return 0;
  k_sample<<<(uint32_t)((count + kThreads - 1) / kThreads), kThreads, 0, s>>>(img, cube, byte_offset, uvw, out,
                                                                               count);
  return 1;
}

int launch_tiles_pack(const uint32_t *color, uint32_t width, uint32_t height, uint32_t rank, uint32_t world,
                      uint32_t *dst, cudaStream_t s)
{
  const uint32_t tiles_x = (width + VB200_TILE - 1) / VB200_TILE, tiles_y = (height + VB200_TILE - 1) / VB200_TILE;
  const uint32_t ntiles = tiles_x * tiles_y;
  const uint32_t slots = (ntiles + world - 1) / world;
  k_tiles_pack<<<slots, kThreads, 0, s>>>(color, width, height, tiles_x, ntiles, rank, world, dst);
  return 1;
}

int launch_mgpu_barrier(const Vb200PeerSet &flags, uint32_t rank, uint32_t world, uint32_t epoch, uint32_t *timed_out,
                        cudaStream_t s)
{
  k_mgpu_barrier<<<1, 32, 0, s>>>(flags, rank, world, epoch, timed_out);
  return 1;
}

int launch_mgpu_push(const Vb200PeerSet &bufs, void *multicast, uint32_t rank, uint32_t world, uint64_t offset,
                     uint64_t bytes, cudaStream_t s)
{
  if(!bytes)
    return 0;
  // callers pass 16-byte aligned ranges (slices of mirrors are cut at multiples of 16)
  k_mgpu_push<<<grid_for(bytes / 16, 4), kThreads, 0, s>>>(bufs, (uint4 *)multicast, rank, world, offset / 16, bytes / 16);
  return 1;
}

int launch_tiles_unpack(uint32_t *color, uint32_t width, uint32_t height, uint32_t world, uint32_t slots_per_rank,
                        const uint32_t *src, cudaStream_t s)
{
  const uint32_t tiles_x = (width + VB200_TILE - 1) / VB200_TILE, tiles_y = (height + VB200_TILE - 1) / VB200_TILE;
  const uint32_t ntiles = tiles_x * tiles_y;
  k_tiles_unpack<<<ntiles, kThreads, 0, s>>>(color, width, height, tiles_x, ntiles, world, slots_per_rank, src);
  return 1;
}
}    // namespace vb200
