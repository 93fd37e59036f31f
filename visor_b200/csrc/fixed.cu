// fixed.cu — the shader-independent sm_100a kernels of the draw path.
//
//   K0  clear                     ClearTarget                      rasterizer.cpp:312-361
//   K2  triangle assembly+setup   GetIndex, ShadeVerts assembly,   rasterizer.cpp:100-232, 385-452
//                                 DrawTriangles setup / cull / bbox
//   K3  binning                   the per-triangle 32x32 block      rasterizer.cpp:454-515
//                                 split + FIFO queue, restated as
//                                 per-screen-tile ordered lists:
//                                 count -> exclusive scan -> fill -> per-tile sort by triangle id
//   K5' stand-alone sampler       sample_tex_wrapped/_cube_wrapped texture_sampling.cpp:139-250
//   K6  tile pack / unpack        (sort-first multi-GPU; no reference equivalent)
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
// K2 + K3 count/fill share the triangle -> tile-range traversal.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tile_owned(uint32_t tile, uint32_t rank, uint32_t world)
{
  return world <= 1u || (tile % world) == rank;
}

// Visit every owned tile of the inclusive range packed in `tiles`, handing `f` a group of lanes that
// target the SAME tile: f(tile, tri, rank_in_group, group_size, is_leader). Neighbouring triangles of a
// mesh land in the same tile, so per-lane atomics on the tile counters would serialise in the L2
// atomic unit; __match_any_sync lets one lane per (warp, tile) do the atomic for the whole group.
// Ranges of up to 2x2 tiles (every triangle smaller than a tile) take the matched path; larger ones
// (full-screen triangles cover thousands of tiles) are spread over the whole CTA, one tile per thread.
// Must be called by every thread of the CTA (it synchronises).
template <typename F>
__device__ __forceinline__ void for_each_tile(uint32_t tiles, bool alive, uint32_t tiles_x, uint32_t rank,
                                              uint32_t world, uint32_t tri, F f)
{
  const uint32_t tx0 = tiles & 0xffu, ty0 = (tiles >> 8) & 0xffu, tx1 = (tiles >> 16) & 0xffu, ty1 = tiles >> 24;
  const uint32_t nx = alive ? (tx1 - tx0 + 1u) : 0u, ny = alive ? (ty1 - ty0 + 1u) : 0u;
  const bool big = nx > 2u || ny > 2u;
  const uint32_t lane = threadIdx.x & 31u;
  if(__any_sync(0xffffffffu, alive && !big))
  {
#pragma unroll
    for(uint32_t q = 0; q < 4u; q++)
    {
      const uint32_t qx = q & 1u, qy = q >> 1;
      uint32_t tile = 0xffffffffu;
      if(alive && !big && qx < nx && qy < ny)
      {
        tile = (ty0 + qy) * tiles_x + tx0 + qx;
        if(!tile_owned(tile, rank, world))
          tile = 0xffffffffu;
      }
      if(!__any_sync(0xffffffffu, tile != 0xffffffffu))
        continue;    // most triangles touch one tile: the other three quadrants are empty for the whole warp
      const uint32_t peers = __match_any_sync(0xffffffffu, tile);
      if(tile != 0xffffffffu)
        f(tile, tri, __popc(peers & ((1u << lane) - 1u)), __popc(peers), (uint32_t)(__ffs(peers) - 1) == lane, peers);
    }
  }
  // Large ranges: queued in shared memory, then every thread of the CTA takes tiles of each queued
  // triangle (one warp walking the ~1000 tiles of a cube face alone made the fill pass of a 12-triangle
  // draw take 36 us: one returning atomic per tile, 32 at a time).
  __shared__ uint32_t s_big[kThreads][2];
  __shared__ uint32_t s_nbig;
  if(threadIdx.x == 0)
    s_nbig = 0;
  __syncthreads();
  if(alive && big)
  {
    const uint32_t slot = atomicAdd(&s_nbig, 1u);
    s_big[slot][0] = tiles;
    s_big[slot][1] = tri;
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
      const uint32_t tile = (b_ty0 + k / b_nx) * tiles_x + b_tx0 + k % b_nx;
      if(tile_owned(tile, rank, world))
        f(tile, b_tri, 0u, 1u, true, 0u);    // distinct tiles per thread: every thread is its own group
    }
  }
}

// Each thread sets up kSetupPerThread triangles, phase by phase, so that the index loads of all of them
// and then the vertex gathers of all of them are in flight together (the kernel is a chain of two
// dependent loads per triangle and little else).
constexpr int kSetupPerThread = 2;

__global__ void __launch_bounds__(kThreads) k_setup(const Vb200SetupParams p)
{
  uint32_t t[kSetupPerThread], s0[kSetupPerThread], s1[kSetupPerThread], s2[kSetupPerThread];
  uint32_t tiles[kSetupPerThread];
  bool alive[kSetupPerThread];
  float invarea[kSetupPerThread];
  int2 va[kSetupPerThread], vb[kSetupPerThread], vc[kSetupPerThread];

  // ---- 1. triangle assembly (rasterizer.cpp:128-232): list = (3t, 3t+1, 3t+2); strip alternates
  // (t, t+1, t+2) / (t+1, t, t+2) to preserve winding; GetIndex (:100-119)
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
  {
    t[k] = (blockIdx.x * kSetupPerThread + k) * blockDim.x + threadIdx.x;
    alive[k] = t[k] < p.num_tris;
    tiles[k] = 0xffffffffu;
    invarea[k] = 0.0f;
    s0[k] = s1[k] = s2[k] = 0u;
    if(alive[k])
    {
      uint32_t c0, c1, c2;
      if(p.topology == 3u)
      {
        c0 = p.first + 3u * t[k];
        c1 = c0 + 1u;
        c2 = c0 + 2u;
      }
      else
      {
        const uint32_t b = p.first + t[k];
        c0 = (t[k] & 1u) ? b + 1u : b;
        c1 = (t[k] & 1u) ? b : b + 1u;
        c2 = b + 2u;
      }
      if(p.indexed)
      {
        c0 = load_index(p.ib, p.index_type, c0);
        c1 = load_index(p.ib, p.index_type, c1);
        c2 = load_index(p.ib, p.index_type, c2);
      }
      s0[k] = c0;
      s1[k] = c1;
      s2[k] = c2;
    }
  }
  const uint32_t base = (p.indexed && p.range) ? p.range[0] : p.base_vertex;
  // ---- 2. window positions of the three corners (the raster record's first 8 bytes)
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
  {
    s0[k] -= base;
    s1[k] -= base;
    s2[k] -= base;
    alive[k] = alive[k] && s0[k] < p.capacity && s1[k] < p.capacity && s2[k] < p.capacity;
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
      invarea[k] = __fdiv_rn(1.0f, (float)(area2 < 0 ? -area2 : area2));    // rasterizer.cpp:448
      const int flipped = (p.front_face == 1u) ? -area2 : area2;
      if(area2 == 0)
        alive[k] = false;
      else if(flipped > 0 && (p.cull_mode & 1u))
        alive[k] = false;
      else if(flipped < 0 && (p.cull_mode & 2u))
        alive[k] = false;
    }
    // statistics: one atomic per CTA on a spread counter (a per-warp atomic on ONE word costs ~20 us per
    // million triangles: same-address atomics serialise in L2)
    survivors += (uint32_t)__syncthreads_count(alive[k]);
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
    if(t[k] < p.num_tris)
    {
      *(int4 *)(p.tri + t[k]) = make_int4((int)s0[k], (int)s1[k], (int)s2[k], __float_as_int(invarea[k]));
      p.tri_tiles[t[k]] = alive[k] ? tiles[k] : 0xffffffffu;
    }
  }
  if(threadIdx.x == 0 && survivors)
    atomicAdd(&p.counters->slot[blockIdx.x & (VB200_COUNTER_SLOTS - 1)].triangles_out, (unsigned long long)survivors);
  uint32_t *cnt = p.tile_count;
#pragma unroll
  for(int k = 0; k < kSetupPerThread; k++)
    for_each_tile(tiles[k], alive[k], p.tiles_x, p.owner_rank, p.owner_world, t[k],
                  [cnt](uint32_t tile, uint32_t, uint32_t, uint32_t group, bool leader, uint32_t) {
                    if(leader)
                      atomicAdd(&cnt[tile], group);
                  });
}

// exclusive scan of the per-tile counts (<= 65536 tiles) by one CTA; also points the fill cursors at the offsets.
// Each thread owns 4*V consecutive counters held in registers (16-byte loads/stores; the arrays are
// allocated with padding to a multiple of 4096 entries), so the kernel is one load, one block scan,
// one store.
template <int V>
__device__ __forceinline__ void scan_body(const uint32_t *tile_count, uint32_t *tile_offset, uint32_t *tile_cursor,
                                          uint32_t ntiles, uint32_t *total, volatile unsigned long long *host_total,
                                          uint32_t seq)
{
  __shared__ uint32_t warp_sums[32];
  const uint32_t begin = threadIdx.x * 4u * V;
  uint4 c[V];
  uint32_t sum = 0;
#pragma unroll
  for(int v = 0; v < V; v++)
  {
    c[v] = (begin + 4u * v < ntiles) ? *(const uint4 *)(tile_count + begin + 4u * v) : make_uint4(0, 0, 0, 0);
    // entries past ntiles inside the last vector are padding and hold 0
    sum += c[v].x + c[v].y + c[v].z + c[v].w;
  }
  uint32_t incl = sum;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if(lane >= (uint32_t)o)
      incl += v;
  }
  if(lane == 31u)
    warp_sums[warp] = incl;
  __syncthreads();
  if(warp == 0)
  {
    const uint32_t w = warp_sums[lane];
    uint32_t wi = w;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
      if(lane >= (uint32_t)o)
        wi += v;
    }
    warp_sums[lane] = wi - w;    // exclusive
    if(lane == 31u)
    {
      *total = wi;
      if(host_total)
      {
        // mapped pinned host word the host polls instead of synchronising the stream. A posted write,
        // deliberately not fenced: the CTA would otherwise sit out a PCIe round trip before its final
        // stores; the host only needs the value eventually (it is checking for list overflow)
        *host_total = ((unsigned long long)seq << 32) | wi;
      }
    }
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + incl - sum;
#pragma unroll
  for(int v = 0; v < V; v++)
  {
    if(begin + 4u * v >= ntiles)
      break;
    uint4 o;
    o.x = run;
    o.y = o.x + c[v].x;
    o.z = o.y + c[v].y;
    o.w = o.z + c[v].z;
    run = o.w + c[v].w;
    *(uint4 *)(tile_offset + begin + 4u * v) = o;
    *(uint4 *)(tile_cursor + begin + 4u * v) = o;    // the fill pass appends at cursor++ (starts at the offset)
  }
}

__global__ void __launch_bounds__(1024) k_scan(const uint32_t *tile_count, uint32_t *tile_offset,
                                              uint32_t *tile_cursor, uint32_t ntiles, uint32_t *total,
                                              volatile unsigned long long *host_total, uint32_t seq)
{
  if(ntiles <= 4096u)
    scan_body<1>(tile_count, tile_offset, tile_cursor, ntiles, total, host_total, seq);
  else if(ntiles <= 8192u)
    scan_body<2>(tile_count, tile_offset, tile_cursor, ntiles, total, host_total, seq);
  else if(ntiles <= 16384u)
    scan_body<4>(tile_count, tile_offset, tile_cursor, ntiles, total, host_total, seq);
  else if(ntiles <= 32768u)
    scan_body<8>(tile_count, tile_offset, tile_cursor, ntiles, total, host_total, seq);
  else
    scan_body<16>(tile_count, tile_offset, tile_cursor, ntiles, total, host_total, seq);
}

// The fill pass is bound by the round trip of its returning atomics (cursor += group), so each thread
// appends kFillPerThread triangles and keeps that many atomics in flight: the atomics of all of them are
// issued before the first result is used.
constexpr int kFillPerThread = 2;

__global__ void __launch_bounds__(kThreads) k_fill(const Vb200SetupParams p, const uint32_t *tile_offset,
                                                  uint32_t *tile_cursor, uint32_t *list, uint32_t capacity,
                                                  const uint32_t *total)
{
  // speculative launch: the host sized `list` before the pair total was known; if it does not fit,
  // do nothing (the tile kernel does the same) and let the host retry with a larger list
  if(*total > capacity)
    return;
  const uint32_t lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
  uint32_t tri[kFillPerThread], tiles[kFillPerThread];
  bool alive[kFillPerThread];
#pragma unroll
  for(int k = 0; k < kFillPerThread; k++)
  {
    tri[k] = (blockIdx.x * kFillPerThread + k) * blockDim.x + threadIdx.x;
    tiles[k] = tri[k] < p.num_tris ? __ldg(p.tri_tiles + tri[k]) : 0xffffffffu;
    alive[k] = tiles[k] != 0xffffffffu;
  }
  // ranges of up to 2x2 tiles, quadrant by quadrant (same grouping as for_each_tile's matched path)
#pragma unroll
  for(uint32_t q = 0; q < 4u; q++)
  {
    const uint32_t qx = q & 1u, qy = q >> 1;
    uint32_t tile[kFillPerThread], peers[kFillPerThread], base[kFillPerThread];
    bool any = false;
#pragma unroll
    for(int k = 0; k < kFillPerThread; k++)
    {
      const uint32_t tx0 = tiles[k] & 0xffu, ty0 = (tiles[k] >> 8) & 0xffu;
      const uint32_t nx = ((tiles[k] >> 16) & 0xffu) - tx0 + 1u, ny = (tiles[k] >> 24) - ty0 + 1u;
      tile[k] = 0xffffffffu;
      if(alive[k] && nx <= 2u && ny <= 2u && qx < nx && qy < ny)
      {
        tile[k] = (ty0 + qy) * p.tiles_x + tx0 + qx;
        if(!tile_owned(tile[k], p.owner_rank, p.owner_world))
          tile[k] = 0xffffffffu;
      }
      any |= tile[k] != 0xffffffffu;
    }
    if(!__any_sync(0xffffffffu, any))
      continue;
#pragma unroll
    for(int k = 0; k < kFillPerThread; k++)
    {
      peers[k] = __match_any_sync(0xffffffffu, tile[k]);
      base[k] = 0;
      if(tile[k] != 0xffffffffu && (uint32_t)(__ffs(peers[k]) - 1) == lane)
        base[k] = atomicAdd(&tile_cursor[tile[k]], __popc(peers[k]));    // cursors start at the CSR offsets
    }
#pragma unroll
    for(int k = 0; k < kFillPerThread; k++)
    {
      base[k] = __shfl_sync(0xffffffffu, base[k], __ffs(peers[k]) - 1);
      const uint32_t pos = base[k] + __popc(peers[k] & below);
      if(tile[k] != 0xffffffffu && pos < capacity)
        list[pos] = tri[k];
    }
  }
  // larger ranges: the CTA-wide walk of for_each_tile (its matched path finds nothing to do here)
#pragma unroll
  for(int k = 0; k < kFillPerThread; k++)
  {
    const uint32_t nx = ((tiles[k] >> 16) & 0xffu) - (tiles[k] & 0xffu) + 1u;
    const uint32_t ny = (tiles[k] >> 24) - ((tiles[k] >> 8) & 0xffu) + 1u;
    const bool big = alive[k] && (nx > 2u || ny > 2u);
    for_each_tile(tiles[k], big, p.tiles_x, p.owner_rank, p.owner_world, tri[k],
                  [=](uint32_t tile, uint32_t t, uint32_t, uint32_t, bool, uint32_t) {
                    const uint32_t pos = atomicAdd(&tile_cursor[tile], 1u);
                    if(pos < capacity)
                      list[pos] = t;
                  });
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

__global__ void __launch_bounds__(kThreads) k_sort(uint32_t *list, const uint32_t *tile_offset,
                                                  const uint32_t *tile_count, const uint32_t *total,
                                                  uint32_t capacity)
{
  __shared__ __align__(16) uint32_t s[kSortSmem];
  if(*total > capacity)
    return;
  const uint32_t tile = blockIdx.x;
  const uint32_t n = tile_count[tile];
  if(n < 2u)
    return;
  uint32_t *a = list + tile_offset[tile];
  int unsorted = 0;
  for(uint32_t i = threadIdx.x; i + 1u < n; i += blockDim.x)
    unsorted |= a[i] > a[i + 1u];
  if(!__syncthreads_or(unsorted))
    return;
  if(n <= 512u)
  {
    // short lists (the common case: a few hundred triangles per tile): rank sort. Triangle ids are
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
  const uint32_t per_cta = kThreads * kSetupPerThread;
  k_setup<<<(p.num_tris + per_cta - 1) / per_cta, kThreads, 0, s>>>(p);
  return 1;
}

int launch_scan(const uint32_t *tile_count, uint32_t *tile_offset, uint32_t *tile_cursor, uint32_t ntiles,
                uint32_t *total, unsigned long long *host_total_dev, uint32_t seq, cudaStream_t s)
{
  k_scan<<<1, 1024, 0, s>>>(tile_count, tile_offset, tile_cursor, ntiles, total, host_total_dev, seq);
  return 1;
}

int launch_fill(const Vb200SetupParams &p, const uint32_t *tile_offset, uint32_t *tile_cursor, uint32_t *list,
                uint32_t capacity, const uint32_t *total, cudaStream_t s)
{
  if(!p.num_tris)
    return 0;
  const uint32_t per_cta = kThreads * kFillPerThread;
  k_fill<<<(p.num_tris + per_cta - 1) / per_cta, kThreads, 0, s>>>(p, tile_offset, tile_cursor, list, capacity, total);
  return 1;
}

int launch_sort(uint32_t *list, const uint32_t *tile_offset, const uint32_t *tile_count, uint32_t ntiles,
                const uint32_t *total, uint32_t capacity, cudaStream_t s)
{
  k_sort<<<ntiles, kThreads, 0, s>>>(list, tile_offset, tile_count, total, capacity);
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

int launch_tiles_unpack(uint32_t *color, uint32_t width, uint32_t height, uint32_t world, uint32_t slots_per_rank,
                        const uint32_t *src, cudaStream_t s)
{
  const uint32_t tiles_x = (width + VB200_TILE - 1) / VB200_TILE, tiles_y = (height + VB200_TILE - 1) / VB200_TILE;
  const uint32_t ntiles = tiles_x * tiles_y;
  k_tiles_unpack<<<ntiles, kThreads, 0, s>>>(color, width, height, tiles_x, ntiles, world, slots_per_rank, src);
  return 1;
}
}    // namespace vb200
