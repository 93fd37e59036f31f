// scaffold.cu — hand-written sm_100a kernel scaffolds that call the JIT-lowered shader functions.
//
// Compiled ONCE at build time to PTX (nvcc -ptx, -fmad=false) and embedded in the library; at the first
// draw that needs a kernel the runtime splices the PTX of the vertex or fragment entry point it calls
// (spirv_ptx.cpp: vb200_vs / vb200_fs) into that kernel's text and compiles the module whole with ptxas
// (runtime.cpp: compileKernel), so the shader is inlined.  Replaces:
//   ShadeVerts + ToWindow                 rasterizer.cpp:121-253          -> vb200_k_vertex
//   ProcessTriangles (+FS call, blend)    rasterizer.cpp:522-696          -> vb200_k_tile_ordered
//   GetVertexAttributeData                spirv_compile.cpp:572-627       -> vb200_fetch_attr
//   sample_tex_wrapped / sample_cube_wrapped / CacheCoord texel fetch
//                                         texture_sampling.cpp:34-250     -> vb200_sample_tex/_cube
//
// Arithmetic contract: every float op below that feeds coverage, depth or
// colour is an explicit round-to-nearest, unfused intrinsic (__fmul_rn, __fadd_rn, __fdiv_rn ...) in
// the reference's evaluation order; int() is x86 cvttss2si (INT_MIN on overflow/NaN).
#include <cuda_runtime.h>
#include <stdint.h>
#include "device_types.h"
#include "kernels.h"
#include "raster_common.cuh"

// tuning switches (see `make variants`)
#define VB200_PRAGMA(x) _Pragma(#x)
#define VB200_UNROLL(n) VB200_PRAGMA(unroll n)
#ifndef VB200_RESOLVE_THREADS
// CTA size of the resolve kernels = triangles set up per round. A tile of a dense mesh holds ~100 triangles
// and ~10 steps of the row stream: four warps per tile keep the set-up lanes busy and the spread between the
// first and the last warp to finish the stream small (with eight, a third of all warp time was barrier wait).
#define VB200_RESOLVE_THREADS 128
#endif
#ifndef VB200_TICKETS
#define VB200_TICKETS 0    // resolve kernels: dynamic hand-out of the row stream's steps (measured: no gain over the fixed split)
#endif
#ifndef VB200_PB_UNROLL
#define VB200_PB_UNROLL 1
#endif

extern "C" __device__ float4 vb200_vs(const Vb200Env *env, unsigned vid, float4 *interps_out);
// The fragment entry point's result: the colour and, for shaders with OpKill (extended mode), whether the
// invocation was discarded. Returned by value: after ptxas has inlined the shader both are plain registers, and
// for a shader without OpKill `killed` is the constant 0 (the tests on it fold away).
struct __align__(16) Vb200FsOut
{
  float4 color;
  uint32_t killed;
  uint32_t pad[3];
};
extern "C" __device__ Vb200FsOut vb200_fs(const Vb200Env *env, float b0, float b1, float b2, const float4 *v0,
                                          const float4 *v1, const float4 *v2);

// ------------------------------------------------------------------------------------------------
// GetVertexAttributeData (spirv_compile.cpp:572-627)
// ------------------------------------------------------------------------------------------------
extern "C" __device__ float4 vb200_fetch_attr(const Vb200Env *env, unsigned attr, unsigned vid)
{
  const Vb200Attr a = env->attrs[attr & 15];
  // byte *ptr = vb.bytes + vb.offset; ptr += attr.offset; ptr += attr.stride * vertexIndex (32-bit product)
  const uint8_t *ptr = env->vb[a.vb & 3] + a.offset + (uint32_t)(a.stride * vid);
  float4 out = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
  const uintptr_t addr = (uintptr_t)ptr;
  switch(a.format)
  {
    case 109: case 107: case 108:    // R32G32B32A32_{SFLOAT,UINT,SINT}: raw 32-bit lanes
      if((addr & 15) == 0)
        out = __ldg((const float4 *)ptr);
      else if((addr & 3) == 0)
      {
        const float *f = (const float *)ptr;
        out = make_float4(__ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3));
      }
      else
        out = make_float4(vb200_ld_f32_unaligned(ptr), vb200_ld_f32_unaligned(ptr + 4),
                          vb200_ld_f32_unaligned(ptr + 8), vb200_ld_f32_unaligned(ptr + 12));
      break;
    case 106: case 104: case 105:    // R32G32B32
      if((addr & 7) == 0)
      {
        float2 xy = __ldg((const float2 *)ptr);
        out.x = xy.x;
        out.y = xy.y;
        out.z = __ldg((const float *)ptr + 2);
      }
      else if((addr & 3) == 0)
      {
        const float *f = (const float *)ptr;
        out.x = __ldg(f);
        out.y = __ldg(f + 1);
        out.z = __ldg(f + 2);
      }
      else
      {
        out.x = vb200_ld_f32_unaligned(ptr);
        out.y = vb200_ld_f32_unaligned(ptr + 4);
        out.z = vb200_ld_f32_unaligned(ptr + 8);
      }
      break;
    case 103: case 101: case 102:    // R32G32
      if((addr & 7) == 0)
      {
        float2 xy = __ldg((const float2 *)ptr);
        out.x = xy.x;
        out.y = xy.y;
      }
      else
      {
        out.x = vb200_ld_f32_unaligned(ptr);
        out.y = vb200_ld_f32_unaligned(ptr + 4);
      }
      break;
    case 100: case 98: case 99:    // R32
      out.x = vb200_ld_f32_unaligned(ptr);
      break;
    case 37:    // R8G8B8A8_UNORM: float(byte) / 255.0f
    {
      uint32_t u = __float_as_uint(vb200_ld_f32_unaligned(ptr));
      out.x = __fdiv_rn((float)((u & 0x000000ffu) >> 0), 255.0f);
      out.y = __fdiv_rn((float)((u & 0x0000ff00u) >> 8), 255.0f);
      out.z = __fdiv_rn((float)((u & 0x00ff0000u) >> 16), 255.0f);
      out.w = __fdiv_rn((float)((u & 0xff000000u) >> 24), 255.0f);
      break;
    }
    default: break;    // the reference asserts; the host rejects such pipelines before launch
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// texture unit (texture_sampling.cpp) — shared with the stand-alone sampler kernel in fixed.cu
// ------------------------------------------------------------------------------------------------
// float(byte) / 255.0f for every byte value (texture_sampling.cpp:121-133, rasterizer.cpp:595-599), filled by
// each tile kernel before its first barrier: texel conversion and the blend's destination read become
// shared-memory loads instead of IEEE divisions (sixteen per bilinear sample).
#if !VB200_UNORM_NEWTON
__shared__ float vb200_s_unorm[256];
#else
#define vb200_s_unorm ((const float *)nullptr)    // the Newton variant of vb200_unorm8 never reads the table
#endif

extern "C" __device__ float4 vb200_sample_tex(float u, float v, const Vb200Image *img, unsigned long long byteOffs)
{
  return vb200_sample_tex_impl(u, v, img, byteOffs, vb200_s_unorm);
}
extern "C" __device__ float4 vb200_sample_cube(float x, float y, float z, const Vb200Image *img)
{
  return vb200_sample_cube_impl(x, y, z, img, vb200_s_unorm);
}

// ------------------------------------------------------------------------------------------------
// K1: vertex fetch + VS + window transform, one thread per UNIQUE vertex.
// The reference re-shades every index (rasterizer.cpp:136-148); the VS is a pure function of the
// index value so shading each referenced vertex once is result-identical.
// ------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(128) vb200_k_vertex(const __grid_constant__ Vb200Env env,
                                                               const Vb200VertexParams p)
{
  asm volatile("griddepcontrol.launch_dependents;");    // the setup kernel may be scheduled as this grid drains
  for(uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < p.tile_count_n; j += gridDim.x * blockDim.x)
    p.tile_count[j] = 0u;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  Vb200VertexSpan sp = p.span0;
  uint32_t count = p.count;
  if(p.range)
  {
    // the lone draw whose span is measured on the device (k_index_range): vertices [min, max], never beyond
    // the bound vertex buffers (a malformed index must not make the fetch run past them)
    const uint32_t lo = p.range[0], hi = p.range[1];
    if(hi < lo || lo >= p.vertex_bound)
      return;
    sp.src_base = lo;
    count = min(min(hi - lo + 1u, count), p.vertex_bound - lo);
  }
  if(i >= count)
    return;
  if(p.spans)
  {
    uint32_t lo = 0, hi = p.num_spans;    // the last span whose slot_base is <= i (empty spans never match)
    while(hi - lo > 1u)
    {
      const uint32_t mid = (lo + hi) >> 1;
      if(__ldg(&p.spans[mid].slot_base) <= i)
        lo = mid;
      else
        hi = mid;
    }
    sp = p.spans[lo];
  }
  const float4 pos = vb200_vs(&env, sp.src_base + (i - sp.slot_base), p.interps + (size_t)i * p.nslots);

  Vb200RasterVertex rv;
  // ToWindow (rasterizer.cpp:248-249):
  //   win.x = int((x / w + 1.0f) * 0.5f * W);  win.y = int((y * -1.0f / w + 1.0f) * 0.5f * H)
  rv.x = vb200_cvtt(__fmul_rn(__fmul_rn(__fadd_rn(__fdiv_rn(pos.x, pos.w), 1.0f), 0.5f), (float)p.width));
  rv.y = vb200_cvtt(
      __fmul_rn(__fmul_rn(__fadd_rn(__fdiv_rn(__fmul_rn(pos.y, -1.0f), pos.w), 1.0f), 0.5f), (float)p.height));
  // per-vertex part of setup (rasterizer.cpp:449-452): invw = 1/w, depth = z * invw
  rv.invw = __frcp_rn(pos.w);    // 1.0f / w, correctly rounded: the same bits as the general division
  rv.depth = __fmul_rn(pos.z, rv.invw);
  p.rv[i] = rv;
}

// ------------------------------------------------------------------------------------------------
// K4 (ordered): one CTA per 32x32 screen tile; exact for ANY state because every pixel sees its
// fragments in submission order (blending, depth ties, EQUAL/NOT_EQUAL, test-without-write).
//
// Warp-autonomous: warp w owns the 16x8 pixel region at ((w&1)*16, (w>>1)*8) of the tile and walks the
// tile's (sorted) triangle list by itself — no CTA barrier inside the loop, so a warp whose region is
// busy never stalls the others.
//   1. 32 triangles at a time: each lane loads one setup record, derives the edge coefficients and tests
//      the bbox against the warp's region; a ballot leaves only the triangles that touch the region.
//   2. per surviving triangle, in list order: every lane tests its 4 pixels (integer edge functions),
//      the covered pixels of the whole region are ranked with ballots + popc and their indices are
//      compacted into a per-warp queue;
//   3. the queue is consumed 32 fragments per pass, all lanes busy: depth test, perspective, the PTX
//      fragment function, blend, truncating BGR pack. Colour/depth of the region live in shared memory
//      so any lane can shade any pixel; they are read from HBM once and written back once.
// ------------------------------------------------------------------------------------------------
struct TriSmem    // 80 B
{
  int A1, B1, C1, A2;
  int B2, C2, area, box;    // box: bbox clipped to the region, region-relative: x0 | y0 << 8 | width << 16 | pixels << 24
  float invarea, invw0, invw1, invw2;
  float d0, d1, d2;
  uint32_t s0;
  uint32_t s1, s2, wrecip, pad1;    // wrecip = ceil(65536 / width): p / width == (p * wrecip) >> 16 for p < 4096
};

// blend factor (rasterizer.cpp:601-653) as selects on the uniform factor enum
__device__ __forceinline__ float vb200_factor_sel(uint32_t f, float alpha, float oneMinusAlpha)
{
  return f == 6u ? alpha : (f == 7u ? oneMinusAlpha : (f == 0u ? 0.0f : 1.0f));
}

// (no min-blocks hint in the source: the runtime adds .minnctapersm at JIT time when the shader at hand
// leaves room for it, see runtime.cpp: compileKernel)
extern "C" __global__ void __launch_bounds__(256) vb200_k_tile_ordered(const __grid_constant__ Vb200Env env,
                                                                     const __grid_constant__ Vb200TileParams p)
{
  __shared__ __align__(16) TriSmem s_tri[8][32];
  __shared__ uint32_t s_col[8][128];
  __shared__ float s_dep[8][128];
  __shared__ uint16_t s_queue[8][256];    // per-warp fragment ring: (triangle slot << 7) | region pixel
  __shared__ uint32_t s_wids[8][64];      // per-warp id queue of the fallback scan (overflowed tile list)

  // sort-first: the grid holds only the tiles this rank owns (tile % world == rank); the others are
  // cleared, drawn and published by their owners
  // (launched with programmatic stream serialization behind the setup / sort kernel: wait for its lists)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tile = blockIdx.x * p.rs.owner_world + p.rs.owner_rank;
  if(tile >= p.rs.tiles_x * p.rs.tiles_y)
    return;
  const uint32_t n = p.tile_count[tile];
  const bool clearColor = (p.clear_flags & 1u) != 0, clearDepth = (p.clear_flags & 2u) != 0;
  if(n == 0 && !p.clear_flags)
    return;
  // More appends than the tile's list holds: the list is incomplete. The warp then finds the tile's
  // triangles by scanning the packed tile ranges of the whole draw, which also yields them in submission order.
  const bool scanList = n > p.list_cap;
  const uint32_t *list = p.list + (size_t)blockIdx.x * p.list_cap;
  uint32_t scanPos = 0, scanQueued = 0;
  const Vb200RasterState &rs = p.rs;
  const uint32_t ty = __umulhi(tile, rs.tiles_x_magic), tx = tile - ty * rs.tiles_x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rx0 = (int)(tx * VB200_TILE) + (warp & 1) * 16, ry0 = (int)(ty * VB200_TILE) + (warp >> 1) * 8;
  const bool depthTest = rs.has_depth && rs.depth_op != 7u;
  const bool depthWrite = rs.has_depth && rs.depth_write;
  const bool blend = rs.blend_enable != 0u && rs.blend_op == 0u;    // only ADD is defined (rasterizer.cpp:657-669)

#if !VB200_UNORM_NEWTON
  vb200_s_unorm[threadIdx.x] = __ldg(p.unorm + threadIdx.x);
#endif
  uint32_t *wcol = s_col[warp];
  float *wdep = s_dep[warp];
  uint16_t *wq = s_queue[warp];
  TriSmem *wtri = s_tri[warp];

  // region pixel i (0..127): x = rx0 + (i & 15), y = ry0 + (i >> 4); lane l loads/tests pixels l + 32j
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    const int i = lane + 32 * j;
    const int x = rx0 + (i & 15), y = ry0 + (i >> 4);
    const bool in = x < (int)rs.width && y < (int)rs.height;
    const size_t idx = (size_t)y * rs.width + x;
    wcol[i] = clearColor ? p.clear_color : (in ? p.color[idx] : 0u);
    wdep[i] = clearDepth ? p.clear_depth : ((in && rs.has_depth) ? p.depth[idx] : 0.0f);
  }
#if !VB200_UNORM_NEWTON
  __syncthreads();    // vb200_s_unorm is shared by all warps; after this the warps run independently
#else
  __syncwarp();    // a warp's region is its own: the warps run independently from the start
#endif
  uint32_t covered = 0, shaded = 0;
  const uint32_t below = (1u << lane) - 1u;
  uint32_t qhead = 0, qcount = 0;              // the warp's fragment ring (uniform across the lanes)

  // One pass over the first `cnt` (<= 32) queued fragments, one per lane. Fragments of DIFFERENT
  // triangles share a pass, so two lanes may hold the same pixel (the two triangles of a quad both
  // cover their common edge). Only the first fragment of each pixel is shaded in this pass; the later
  // ones go back to the front of the ring, in order, and lead the next pass.
  auto shade_pass = [&](uint32_t cnt) {
    const bool mine = (uint32_t)lane < cnt;
    const uint32_t e = wq[(qhead + (uint32_t)lane) & 255u];
    const int i = (int)(e & 127u);
    const uint32_t lanes = __ballot_sync(0xffffffffu, mine);
    bool first = false;
    if(mine)
      first = (__match_any_sync(lanes, i) & below) == 0u;    // no earlier fragment of my pixel in this pass
    const uint32_t later = __ballot_sync(0xffffffffu, mine && !first);
    {
      if(first)
      {
        const TriSmem &t = wtri[e >> 7];
        const int x = i & 15, y = i >> 4;
        const int b1 = t.A1 * x + t.B1 * y + t.C1, b2 = t.A2 * x + t.B2 * y + t.C2, b0 = t.area - (b1 + b2);
        // rasterizer.cpp:552-558
        float n0 = __fmul_rn((float)b0, t.invarea);
        float n1 = __fmul_rn((float)b1, t.invarea);
        float n2 = __fmul_rn((float)b2, t.invarea);
        const float pixdepth = __fadd_rn(__fadd_rn(__fmul_rn(n0, t.d0), __fmul_rn(n1, t.d1)), __fmul_rn(n2, t.d2));
        if(!depthTest || vb200_depth_pass(rs.depth_op, pixdepth, wdep[i]))
        {
          shaded++;
          // perspective correction (rasterizer.cpp:581-588)
          n0 = __fmul_rn(n0, t.invw0);
          n1 = __fmul_rn(n1, t.invw1);
          n2 = __fmul_rn(n2, t.invw2);
          // 1.0f / x, correctly rounded (the dedicated reciprocal: a shorter sequence than the general division,
          // the same bits)
          const float invlen = __frcp_rn(__fadd_rn(__fadd_rn(n0, n1), n2));
          n0 = __fmul_rn(n0, invlen);
          n1 = __fmul_rn(n1, invlen);
          n2 = __fmul_rn(n2, invlen);

          const Vb200FsOut fo = vb200_fs(&env, n0, n1, n2, p.interps + (size_t)t.s0 * rs.nslots,
                                         p.interps + (size_t)t.s1 * rs.nslots, p.interps + (size_t)t.s2 * rs.nslots);
          float4 pix = fo.color;
          const uint32_t cur = wcol[i];
          // a discarded fragment (OpKill, extended mode) leaves colour and depth as they are
          if(fo.killed)
            ;
          else
          {
          if(blend)
          {
            // blend (rasterizer.cpp:593-672): existing = bytes (2,1,0) / 255.0f from the exact table
            const float ex = vb200_unorm8(vb200_s_unorm, (cur >> 16) & 0xffu),
                        ey = vb200_unorm8(vb200_s_unorm, (cur >> 8) & 0xffu), ez = vb200_unorm8(vb200_s_unorm, cur & 0xffu);
            const float oma = __fsub_rn(1.0f, pix.w);
            const float srcF = vb200_factor_sel(rs.src_factor, pix.w, oma);
            const float dstF = vb200_factor_sel(rs.dst_factor, pix.w, oma);
            pix.x = __fadd_rn(__fmul_rn(srcF, pix.x), __fmul_rn(dstF, ex));
            pix.y = __fadd_rn(__fmul_rn(srcF, pix.y), __fmul_rn(dstF, ey));
            pix.z = __fadd_rn(__fmul_rn(srcF, pix.z), __fmul_rn(dstF, ez));
          }
          // truncating BGR store, alpha byte untouched (rasterizer.cpp:674-676)
          const uint32_t r = (uint32_t)__float2int_rz(__fmul_rn(__saturatef(pix.x), 255.0f));
          const uint32_t g = (uint32_t)__float2int_rz(__fmul_rn(__saturatef(pix.y), 255.0f));
          const uint32_t b = (uint32_t)__float2int_rz(__fmul_rn(__saturatef(pix.z), 255.0f));
          wcol[i] = (cur & 0xff000000u) | (r << 16) | (g << 8) | b;
          if(depthWrite)
            wdep[i] = pixdepth;
          }
        }
      }
    }
    const uint32_t keep = __popc(later);
    __syncwarp();    // every lane has consumed its ring entry before the tail of this pass is rewritten
    if(mine && !first)
      wq[(qhead + cnt - keep + __popc(later & below)) & 255u] = (uint16_t)e;
    __syncwarp();    // a later fragment of the same pixel (in a later pass) sees this one; ring updated
    qhead = (qhead + cnt - keep) & 255u;
    qcount -= cnt - keep;
  };

  for(uint32_t base = 0; base < n; base += 32u)
  {
    // ---- 1. load + setup 32 triangles, keep those whose bbox touches this warp's region
    bool touches = false;
    bool have = base + lane < n;
    uint32_t t = 0;
    if(!scanList)
    {
      if(have)
        t = list[base + lane];
    }
    else
    {
      uint32_t *ids = s_wids[warp];
      while(scanQueued < 32u && scanPos < p.num_tris)
      {
        const uint32_t idx = scanPos + (uint32_t)lane;
        const bool match = idx < p.num_tris && vb200_tile_in_range(__ldg(p.tri_tiles + idx), tx, ty);
        const uint32_t mm = __ballot_sync(0xffffffffu, match);
        if(match)
          ids[scanQueued + __popc(mm & below)] = idx;
        scanQueued += __popc(mm);
        scanPos += 32u;
      }
      __syncwarp();
      have = (uint32_t)lane < scanQueued;
      t = have ? ids[lane] : 0u;
      const uint32_t rest = scanQueued - min(scanQueued, 32u);    // < 32: they move to the front
      const uint32_t mv = (uint32_t)lane < rest ? ids[32 + lane] : 0u;
      __syncwarp();
      if((uint32_t)lane < rest)
        ids[lane] = mv;
      __syncwarp();
      scanQueued = rest;
    }
    if(have)
    {
      const Vb200TriSetup su = vb200_load_setup(p, t);
      // MinMax + clamp (rasterizer.cpp:428-435); pixels iterate the half-open box [min, max)
      const int minx = max(0, min(su.x0, min(su.x1, su.x2))), miny = max(0, min(su.y0, min(su.y1, su.y2)));
      const int maxx = min((int)rs.width - 1, max(su.x0, max(su.x1, su.x2)));
      const int maxy = min((int)rs.height - 1, max(su.y0, max(su.y1, su.y2)));
      const int bx0 = max(minx, rx0) - rx0, bx1 = min(maxx, rx0 + 16) - rx0;
      const int by0 = max(miny, ry0) - ry0, by1 = min(maxy, ry0 + 8) - ry0;
      touches = bx0 < bx1 && by0 < by1;
      if(touches)
      {
        // rasterizer.cpp:395-448 — area2, barymul = sign(area2), |area2|; barycentric() (:303-309)
        // with barymul folded in and re-based to region-relative pixel coordinates (exact in the
        // int32 ring): b1 = A1*x + B1*y + C1, b2 = A2*x + B2*y + C2, b0 = |area2| - (b1 + b2)
        const int ABx = su.x1 - su.x0, ABy = su.y1 - su.y0, ACx = su.x2 - su.x0, ACy = su.y2 - su.y0;
        const int area2 = ABx * ACy - ABy * ACx;
        const int sgn = area2 > 0 ? 1 : -1;
        TriSmem s;
        s.A1 = sgn * ACy;
        s.B1 = -sgn * ACx;
        s.C1 = sgn * (ACx * su.y0 - ACy * su.x0) + s.A1 * rx0 + s.B1 * ry0;
        s.A2 = -sgn * ABy;
        s.B2 = sgn * ABx;
        s.C2 = sgn * (ABy * su.x0 - ABx * su.y0) + s.A2 * rx0 + s.B2 * ry0;
        s.area = sgn * area2;
        const int bw = bx1 - bx0;    // 1..16; the box holds at most 16 x 8 = 128 pixels
        s.box = bx0 | (by0 << 8) | (bw << 16) | ((bw * (by1 - by0)) << 24);
        s.invarea = su.invarea;
        s.invw0 = su.invw0;
        s.invw1 = su.invw1;
        s.invw2 = su.invw2;
        s.d0 = su.d0;
        s.d1 = su.d1;
        s.d2 = su.d2;
        s.s0 = su.s0;
        s.s1 = su.s1;
        s.s2 = su.s2;
        s.wrecip = (65536u + (uint32_t)bw - 1u) / (uint32_t)bw;
        s.pad1 = 0;
        wtri[lane] = s;
      }
    }
    uint32_t hits = __ballot_sync(0xffffffffu, touches);
    __syncwarp();

    // ---- 2. surviving triangles in list order: coverage -> fragment ring
    while(hits)
    {
      const int k = __ffs(hits) - 1;
      hits &= hits - 1u;
      const TriSmem &t = wtri[k];
      // The pixels of the triangle's clipped bbox (row-major, at most 128), 32 per pass, one per lane: a
      // particle-sized box takes one or two ballots where testing the whole 16x8 region always took four.
      const int bx0 = t.box & 0xff, by0 = (t.box >> 8) & 0xff, bw = (t.box >> 16) & 0xff;
      const uint32_t npx = (uint32_t)t.box >> 24;
      const uint32_t tag = (uint32_t)k << 7;
      uint32_t tail = qhead + qcount, total = 0;
#pragma unroll
      for(uint32_t j = 0; j < 4u; j++)
      {
        if(32u * j >= npx)
          break;
        const uint32_t pi = (uint32_t)lane + 32u * j;
        const int dy = (int)((pi * t.wrecip) >> 16), dx = (int)pi - dy * bw;
        const int x = bx0 + dx, y = by0 + dy;
        const int b1 = t.A1 * x + t.B1 * y + t.C1, b2 = t.A2 * x + t.B2 * y + t.C2;
        const int b0 = t.area - (b1 + b2);
        // covered iff all three >= 0 (rasterizer.cpp:549)
        const bool inside = pi < npx && ((b0 | b1 | b2) >= 0);
        const uint32_t m = __ballot_sync(0xffffffffu, inside);
        // append the covered pixels to the ring, tagged with the triangle's slot
        if(inside)
          wq[(tail + __popc(m & below)) & 255u] = (uint16_t)(tag | (uint32_t)(y * 16 + x));
        tail += __popc(m);
        total += __popc(m);
      }
      if(total == 0u)
        continue;
      covered += (lane == 0) ? total : 0u;
      qcount += total;
      __syncwarp();
      // ---- 3. full 32-fragment passes as soon as the ring holds them
      while(qcount >= 32u)
        shade_pass(32u);
    }
    while(qcount)    // the triangle slots are about to be reused: drain what is left of this batch
      shade_pass(min(qcount, 32u));
    __syncwarp();    // wtri is overwritten by the next 32 triangles
  }

#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    const int i = lane + 32 * j;
    const int x = rx0 + (i & 15), y = ry0 + (i >> 4);
    if(x < (int)rs.width && y < (int)rs.height)
    {
      const size_t idx = (size_t)y * rs.width + x;
      vb200_store_color(p, (uint32_t)idx, wcol[i], p.mc_color != nullptr || p.num_peers != 0u);
      if(depthWrite || (clearDepth && rs.has_depth))
        __stcs(p.depth + idx, wdep[i]);
    }
  }
  if(rs.count_fragments)
    vb200_count_fragments(p.counters, covered, shaded);
}

// ------------------------------------------------------------------------------------------------
// K4 (resolve): the same tile decomposition for passes whose per-pixel result does not depend on
// the ORDER fragments arrive in, only on which fragment wins:
//   depth LESS / LESS_OR_EQUAL / GREATER / GREATER_OR_EQUAL with depth write, no blending:
//       the surviving fragment is the min (max) depth; ties go to the first (LESS, GREATER) or last
//       (.._OR_EQUAL) triangle in draw order, exactly what the serial loop of the reference yields
//       (rasterizer.cpp:560-576,680-683: a fragment passes against the depth left by its predecessors);
//   no depth test, or a test against a depth buffer the pass does not modify (write off, or EQUAL):
//       the last passing triangle in draw order wins.
// The supported SPIR-V subset has no side effects or discard, so shading only the winner is exact.
//
// Phase A (triangle-parallel): each thread rasterises whole small triangles of the tile's list
// (bbox <= 64 px in the tile); larger ones are queued in shared memory and swept row by row, one
// 32-pixel row per warp step. Winners are resolved with a 64-bit (depth key, triangle id) atomic
// min in shared memory — no global atomics, lists need no sorting.
// Phase B (pixel-parallel): each thread recomputes the winner's barycentrics bit-exactly, runs the
// fragment stage once per pixel and writes colour/depth rows as full 128-byte lines.
// ------------------------------------------------------------------------------------------------
enum
{
  VB200_RES_MIN_FIRST = 0,    // LESS + write
  VB200_RES_MIN_LAST = 1,     // LESS_OR_EQUAL + write
  VB200_RES_MAX_FIRST = 2,    // GREATER + write
  VB200_RES_MAX_LAST = 3,     // GREATER_OR_EQUAL + write
  VB200_RES_LAST_WINS = 4,    // no test / static test: highest triangle id that passes
};

// Shared-memory accesses of the resolve kernel's inner loop through explicit 32-bit shared-window
// addresses. ptxas otherwise re-derives the window base (S2R SR_CgaCtaId + LEA) inside the loop for
// every access instead of holding it in a register (~8 of 83 instructions per step, plus the S2R latency).
__device__ __forceinline__ uint32_t vb200_smem_addr(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t vb200_lds32(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 vb200_lds128(uint32_t a)
{
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long vb200_lds64(uint32_t a)
{
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long vb200_atoms_cas64(uint32_t a, unsigned long long cmp, unsigned long long val)
{
  unsigned long long old;
  asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val) : "memory");
  return old;
}

// Records of one round (up to 256 triangles of the tile's list, one per thread), staged in shared memory
// as a structure of 16-byte arrays: a gather of one field group by the 8 lanes of a quarter warp touches
// 8 different 16-byte bank groups unless two of them ask for records a multiple of 8 apart (the 96-byte
// array-of-structures records this replaces collided four ways: 9.5 M bank conflicts per C3 frame).
//   e1 = {A1, B1, C1, A2}         b1 = A1*x + B1*y + C1      x, y: pixel inside the tile (0..31)
//   e2 = {B2, C2, |area2|, box}   b2 = A2*x + B2*y + C2,  b0 = |area2| - (b1 + b2)
//                                 box = x0 | y0 << 8 | w << 16 | h << 24: the tile-clipped bbox
//   z  = {1/|area2|, d0, d1, d2}
//   pw = {invw0, invw1, invw2, s0}
//   sv = {s1, s2}
//   key = key id (below)
// (Same int32 ring arithmetic as barycentric(), rasterizer.cpp:303-309, with barymul folded in.)
// key id = triangle index + 1, or ((triangle index + 1) << 8 | record slot) when Vb200RasterState::slot_keys
// is set: ids stay ordered by triangle index, and for a tile whose whole list fits one round (<= 256
// triangles, the common case) phase B reads the winner's record from shared memory instead of gathering
// the triangle and its three vertices from global memory again.

// float -> uint32 whose unsigned order equals the float order (-0 == +0); callers exclude NaN
__device__ __forceinline__ uint32_t vb200_depth_key(float d)
{
  const uint32_t b = __float_as_uint(__fadd_rn(d, 0.0f));    // -0 + 0 = +0; every other value unchanged
  return b ^ ((uint32_t)((int)b >> 31) | 0x80000000u);       // negative: flip all bits; else set the sign bit
}

template <int MODE>
__device__ __forceinline__ unsigned long long vb200_existing_key(float e)
{
  if(MODE == VB200_RES_LAST_WINS)
    return ~0ull;
  // low word when nothing has won the pixel: 0 (.._FIRST: ids are >= 1) or ~0 (.._LAST: ~id is < ~0)
  const uint32_t low = (MODE == VB200_RES_MIN_FIRST || MODE == VB200_RES_MAX_FIRST) ? 0u : 0xffffffffu;
  if(e != e)
    return (unsigned long long)low;    // NaN in the depth buffer: every comparison fails. High word 0 lies
                                       // below every fragment's depth key, so nothing replaces it, and the
                                       // low word still reads "no winner" in phase B
  uint32_t k = vb200_depth_key(e);
  if(MODE == VB200_RES_MAX_FIRST || MODE == VB200_RES_MAX_LAST)
    k = ~k;
  return ((unsigned long long)k << 32) | low;
}

// Visibility key of one covered fragment (barycentric numerators b0..b2). False: it cannot win (NaN depth,
// or it fails the test against the depth the pass does not modify).
template <int MODE>
__device__ __forceinline__ bool vb200_fragment_key(int b0, int b1, int b2, float invarea, float d0, float d1, float d2,
                                                   uint32_t id, bool depthTest, uint32_t depthOp, const float *s_depth,
                                                   int idx, unsigned long long &key)
{
  if(MODE == VB200_RES_LAST_WINS && !depthTest)
  {
    key = (unsigned long long)(~id);
    return true;
  }
  // rasterizer.cpp:552-558
  const float n0 = __fmul_rn((float)b0, invarea);
  const float n1 = __fmul_rn((float)b1, invarea);
  const float n2 = __fmul_rn((float)b2, invarea);
  const float pixdepth = __fadd_rn(__fadd_rn(__fmul_rn(n0, d0), __fmul_rn(n1, d1)), __fmul_rn(n2, d2));
  if(MODE == VB200_RES_LAST_WINS)
  {
    key = (unsigned long long)(~id);
    return vb200_depth_pass(depthOp, pixdepth, s_depth[idx]);
  }
  if(pixdepth != pixdepth)
    return false;    // NaN never passes an ordered comparison
  uint32_t dk = vb200_depth_key(pixdepth);
  if(MODE == VB200_RES_MAX_FIRST || MODE == VB200_RES_MAX_LAST)
    dk = ~dk;
  const uint32_t low = (MODE == VB200_RES_MIN_FIRST || MODE == VB200_RES_MAX_FIRST) ? id : ~id;
  key = ((unsigned long long)dk << 32) | low;
  return true;
}

// Where pixel (x, y) of the tile keeps its visibility key. Rows are rotated against each other by
// 1 << VB200_VIS_SKEW columns: the hits a warp processes together are short runs of rows stacked on top of
// each other (small triangles), and with plain row-major slots the runs of one triangle would all fall into
// the same shared-memory banks (a 256-byte row is a whole number of bank cycles).
#ifndef VB200_VIS_SKEW
#define VB200_VIS_SKEW 2
#endif
__device__ __forceinline__ uint32_t vb200_vis_slot(uint32_t px, uint32_t py)
{
#if VB200_VIS_SKEW
  return (py * VB200_TILE) | ((px + (py << VB200_VIS_SKEW)) & (VB200_TILE - 1u));
#else
  return py * VB200_TILE + px;
#endif
}

// 64-bit min in shared memory has no native atomic: CAS until the slot holds a key <= ours (the common
// case is no attempt at all, or one that succeeds). `seen` is a plain read that may race with other
// warps' CAS on purpose (compute-sanitizer racecheck reports it): keys only ever decrease, so a stale
// value can only cause a CAS attempt that fails and refreshes it.
__device__ __forceinline__ void vb200_vis_min(uint32_t aSlot, unsigned long long seen, unsigned long long key)
{
  while(key < seen)
  {
    const unsigned long long prev = vb200_atoms_cas64(aSlot, seen, key);
    if(prev == seen)
      break;
    seen = prev;
  }
}

template <int MODE>
__device__ __forceinline__ void vb200_tile_resolve_body(const Vb200Env &env, const Vb200TileParams &p)
{
  constexpr int RT = VB200_RESOLVE_THREADS, RW = RT / 32;    // threads / warps of a CTA = triangles per round
  __shared__ unsigned long long vis[VB200_TILE * VB200_TILE];
  __shared__ float s_depth[MODE == VB200_RES_LAST_WINS ? VB200_TILE * VB200_TILE : 1];
  __shared__ int4 s_e1[RT], s_e2[RT], s_z[RT], s_pw[RT];    // records of the current round
  __shared__ int2 s_sv[RT];
  __shared__ uint32_t s_key[RT];
  __shared__ __align__(16) uint32_t s_start[RT + 4];   // first stream row of each record; past the records: stream length
  __shared__ uint2 s_run[RW][64];                    // per warp: the covered runs of the 64 rows of the current step
  __shared__ uint32_t s_wsum[RW];
  __shared__ uint32_t s_ticket;    // next unclaimed step of the row stream
  __shared__ uint32_t s_ids[2 * RT];    // id queue of the fallback scan (overflowed tile list)

  // sort-first: the grid holds only the tiles this rank owns (tile % world == rank); the others are
  // cleared, drawn and published by their owners
  // (launched with programmatic stream serialization behind the setup kernel: wait for its lists)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tile = blockIdx.x * p.rs.owner_world + p.rs.owner_rank;
  if(tile >= p.rs.tiles_x * p.rs.tiles_y)
    return;
  const uint32_t n = p.tile_count[tile];
  const bool clearColor = (p.clear_flags & 1u) != 0, clearDepth = (p.clear_flags & 2u) != 0;
  const Vb200RasterState &rs = p.rs;
  // the tile's whole list fits one round: the records stay staged for phase B
  const bool recordsInSmem = rs.slot_keys && n <= (uint32_t)RT;
  const uint32_t ty = __umulhi(tile, rs.tiles_x_magic), tx = tile - ty * rs.tiles_x;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("" : "+r"(lane), "+r"(warp));    // (kept in registers: ptxas otherwise re-reads %tid inside the loops)
  const int tileX0 = (int)(tx * VB200_TILE), tileY0 = (int)(ty * VB200_TILE);
  if(n == 0)
  {
    // no triangle touches this tile: it only has to receive the folded clears
    if(p.clear_flags)
    {
#pragma unroll
      for(int j = 0; j < VB200_TILE / RW; j++)
      {
        const int x = tileX0 + lane, y = tileY0 + warp + RW * j;
        if(x < (int)rs.width && y < (int)rs.height)
        {
          const uint32_t gi = (uint32_t)y * rs.width + (uint32_t)x;
          if(clearColor)
            vb200_store_color(p, gi, p.clear_color, p.mc_color != nullptr || p.num_peers != 0u);
          if(clearDepth && rs.has_depth)
            __stcs(p.depth + gi, p.clear_depth);
        }
      }
    }
    return;
  }
  // More appends than the tile's list holds: the list is incomplete. The CTA then finds the tile's
  // triangles by scanning the packed tile ranges of the whole draw (exact; needs no host round trip).
  const bool scanList = n > p.list_cap;
  const uint32_t *list = p.list + (size_t)blockIdx.x * p.list_cap;
  uint32_t scanPos = 0, scanQueued = 0;
  const bool depthTest = rs.has_depth && rs.depth_op != 7u;
  const bool depthWrite = rs.has_depth && rs.depth_write;

  // ---- phase A: coverage + visibility, up to 256 triangles of the list per round.
  //  1. one thread per triangle: load, edge setup, tile-clipped bbox -> record in shared memory;
  //  2. the ROW PAIRS of all bboxes are laid end to end into one stream (block-wide exclusive prefix sum of
  //     ceil(height / 2)), cut into 32-unit steps; warp w takes the w-th eighth of the steps, so every warp
  //     does the same amount of work whatever the triangle sizes are. A lane owns two consecutive rows of one
  //     triangle: it walks them with incremental edge functions into coverage masks (five instructions per
  //     candidate pixel, no cross-lane traffic). The covered pixels of a row are one run [xa, xa + L);
  //  3. the warp lays the runs of its 64 rows end to end (warp prefix sum of L) and visits only COVERED
  //     pixels, 32 at a time, every lane busy: depth, key, merge into the pixel's visibility slot.
  // Winners are resolved with a 64-bit (depth key, triangle id) min in shared memory: no global atomics,
  // and the lists need no sorting.
  // Every listed triangle has a non-empty clipped bbox (the binning walked exactly these tile ranges).
  // shared-window addresses of the arrays the inner loops gather from, pinned in registers (left alone, ptxas
  // re-derives each base inside the loops: S2R SR_CgaCtaId + LEA per access)
  uint32_t aVis = vb200_smem_addr(vis), aStart = vb200_smem_addr(s_start), aE1 = vb200_smem_addr(s_e1),
           aE2 = vb200_smem_addr(s_e2), aZ = vb200_smem_addr(s_z), aKey = vb200_smem_addr(s_key);
  asm volatile("" : "+r"(aVis), "+r"(aStart), "+r"(aE1), "+r"(aE2), "+r"(aZ), "+r"(aKey));
  uint32_t covered = 0, shaded = 0;
  for(uint32_t base = 0; base < n; base += RT)
  {
    uint32_t m = min((uint32_t)RT, n - base);    // records in this round
    uint32_t t = 0;
    if(!scanList)
    {
      if(threadIdx.x < m)
        t = list[base + threadIdx.x];
    }
    else
    {
      // fallback: gather the next m ids, in order, from the packed tile ranges
      while(scanQueued < m && scanPos < p.num_tris)
      {
        const uint32_t idx = scanPos + threadIdx.x;
        const bool match = idx < p.num_tris && vb200_tile_in_range(__ldg(p.tri_tiles + idx), tx, ty);
        const uint32_t mm = __ballot_sync(0xffffffffu, match);
        __syncthreads();    // the previous pass is done with s_wsum / s_ids
        if(lane == 0)
          s_wsum[warp] = __popc(mm);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for(int q = 0; q < RW; q++)
        {
          const uint32_t v = s_wsum[q];
          before += (q < warp) ? v : 0u;
          total += v;
        }
        if(match)
          s_ids[scanQueued + before + __popc(mm & ((1u << lane) - 1u))] = idx;
        scanQueued += total;
        scanPos += RT;
      }
      __syncthreads();
      m = min(m, scanQueued);
      if(threadIdx.x < m)
        t = s_ids[threadIdx.x];
      const uint32_t rest = scanQueued - m;    // < RT: they move to the front of the queue
      const uint32_t mv = threadIdx.x < rest ? s_ids[m + threadIdx.x] : 0u;
      __syncthreads();
      if(threadIdx.x < rest)
        s_ids[threadIdx.x] = mv;
      scanQueued = rest;
    }
    const bool have = threadIdx.x < m;
    int4 re1 = make_int4(0, 0, 0, 0), re2 = make_int4(0, 0, 0, (1 << 16) | (1 << 24)), rz = re1, rpw = re1;
    int2 rsv = make_int2(0, 0);
    // the three dependent loads of a triangle (list entry above, record, corners) are in flight while the
    // first round initialises the tile below
    int4 rq = make_int4(0, 0, 0, 0), ra = rq, rb = rq, rc = rq;
    if(have)
    {
      rq = __ldg((const int4 *)(p.tri + t));
      ra = __ldg((const int4 *)(p.rv + (uint32_t)rq.x));
      rb = __ldg((const int4 *)(p.rv + (uint32_t)rq.y));
      rc = __ldg((const int4 *)(p.rv + (uint32_t)rq.z));
    }
    if(base == 0u)
    {
#if !VB200_UNORM_NEWTON
      for(int i = threadIdx.x; i < 256; i += RT)
        vb200_s_unorm[i] = __ldg(p.unorm + i);    // read in phase B, after the barriers below
#endif
      // ---- init: one visibility key per pixel, seeded with the depth already in the buffer
    #pragma unroll
      for(int j = 0; j < VB200_TILE / RW; j++)
      {
        const int ly = warp + RW * j;
        const int x = tileX0 + lane, y = tileY0 + ly;
        const bool in = x < (int)rs.width && y < (int)rs.height;
        float e = 0.0f;
        if(MODE != VB200_RES_LAST_WINS || depthTest)
          e = clearDepth ? p.clear_depth : (in ? p.depth[(size_t)y * rs.width + x] : 0.0f);
        vis[vb200_vis_slot(lane, ly)] = vb200_existing_key<MODE>(e);
        if(MODE == VB200_RES_LAST_WINS)
          s_depth[ly * VB200_TILE + lane] = e;
      }
    }
    if(have)
    {
      const Vb200TriSetup su = vb200_unpack_setup(rq, ra, rb, rc);
      const int ABx = su.x1 - su.x0, ABy = su.y1 - su.y0, ACx = su.x2 - su.x0, ACy = su.y2 - su.y0;
      const int area2 = ABx * ACy - ABy * ACx;
      const int sgn = area2 > 0 ? 1 : -1;
      // MinMax + clamp (rasterizer.cpp:428-435), clipped to the tile; pixels iterate the half-open box
      const int x0 = max(max(0, min(su.x0, min(su.x1, su.x2))), tileX0);
      const int y0 = max(max(0, min(su.y0, min(su.y1, su.y2))), tileY0);
      const int x1 = min(min((int)rs.width - 1, max(su.x0, max(su.x1, su.x2))), tileX0 + VB200_TILE);
      const int y1 = min(min((int)rs.height - 1, max(su.y0, max(su.y1, su.y2))), tileY0 + VB200_TILE);
      const int bw = max(x1 - x0, 1), bh = max(y1 - y0, 1);
      // barycentric() (rasterizer.cpp:303-309) with barymul = sign(area2) folded in, re-based to the tile
      const int A1 = sgn * ACy, B1 = -sgn * ACx, A2 = -sgn * ABy, B2 = sgn * ABx;
      re1 = make_int4(A1, B1, sgn * (ACx * su.y0 - ACy * su.x0) + A1 * tileX0 + B1 * tileY0, A2);
      re2 = make_int4(B2, sgn * (ABy * su.x0 - ABx * su.y0) + A2 * tileX0 + B2 * tileY0, sgn * area2,
                      (x0 - tileX0) | ((y0 - tileY0) << 8) | (bw << 16) | (bh << 24));
      rz = make_int4(__float_as_int(su.invarea), __float_as_int(su.d0), __float_as_int(su.d1), __float_as_int(su.d2));
      rpw = make_int4(__float_as_int(su.invw0), __float_as_int(su.invw1), __float_as_int(su.invw2), (int)su.s0);
      rsv = make_int2((int)su.s1, (int)su.s2);
    }
    // block-wide exclusive scan of the stream units: a unit is a pair of consecutive bbox rows (<= 16 per record)
    const uint32_t mine = have ? (((uint32_t)re2.w >> 24) + 1u) >> 1 : 0u;
    uint32_t incl = mine;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if(lane >= o)
        incl += v;
    }
    __syncthreads();    // previous round (or the init) is done with the records / s_start / s_wsum
    if(lane == 31)
      s_wsum[warp] = incl;
    s_e1[threadIdx.x] = re1;
    s_e2[threadIdx.x] = re2;
    s_z[threadIdx.x] = rz;
    s_pw[threadIdx.x] = rpw;
    s_sv[threadIdx.x] = rsv;
    s_key[threadIdx.x] = rs.slot_keys ? (((t + 1u) << 8) | threadIdx.x) : (t + 1u);
    __syncthreads();
    const uint32_t wsum = lane < RW ? s_wsum[lane] : 0u;
    const uint32_t total = __reduce_add_sync(0xffffffffu, wsum);
    const uint32_t wbase = __reduce_add_sync(0xffffffffu, lane < warp ? wsum : 0u);
    // s_start[r] = first unit of record r in the stream; entries past the round's records hold its length
    s_start[threadIdx.x] = have ? wbase + incl - mine : total;
    if(threadIdx.x == 0)
    {
      s_start[RT] = total;
      s_ticket = 0;
    }
    __syncthreads();

    // Units per step: 32, one per lane — unless the whole stream is shorter than 32 units per warp (a few large
    // triangles: one full step would leave all their pixels, up to 64 rows of 32, to a single warp). Then every
    // warp takes its share of the units, the surplus lanes idle through the row walk, and the hit passes, which
    // are where the time goes, spread over the CTA.
    const uint32_t ups = min(32u, max(1u, (total + (uint32_t)RW - 1u) / (uint32_t)RW));
    const uint32_t steps = ups == 32u ? (total + 31u) >> 5 : (total + ups - 1u) / ups;
    // VB200_TICKETS: steps handed out dynamically (a ticket counter in shared memory) instead of warp w taking
    // the w-th share of them. A step's cost depends on how many pixels its rows cover, so a fixed split lets
    // the warps arrive at the barrier behind the stream up to a step's worth of work apart.
    {
#if VB200_TICKETS
      for(;;)
      {
        uint32_t step = 0;
        if(lane == 0)
          step = atomicAdd(&s_ticket, 1u);
        step = __shfl_sync(0xffffffffu, step, 0);
        if(step >= steps)
          break;
#else
      const uint32_t lastStep = (steps * (uint32_t)(warp + 1)) / RW;
      for(uint32_t step = (steps * (uint32_t)warp) / RW; step < lastStep; step++)
      {
#endif
        const uint32_t k = step * ups;
        const uint32_t g = k + lane;
        // record that owns stream unit k: the last record whose start is <= k. Starts are strictly increasing
        // (every record has at least one unit) and entries past them hold `total` (> k), so it is (number
        // of entries <= k) - 1: every lane counts its share of the entries (independent loads) and one warp
        // reduction adds them up — instead of a dependent binary search.
        uint32_t owner0;
        {
          uint32_t cnt = 0;
#pragma unroll
          for(int q = 0; q < RT / 128; q++)
          {
            const uint4 a = *(const uint4 *)(s_start + lane * (RT / 32) + 4 * q);
            cnt += (a.x <= k) + (a.y <= k) + (a.z <= k) + (a.w <= k);
          }
          owner0 = __reduce_add_sync(0xffffffffu, cnt) - 1u;
        }
        // lane l looks at the start of record owner0+1+l: those that begin inside (k, k+32) split the step
        const uint32_t rel = vb200_lds32(aStart + 4u * min(owner0 + 1u + (uint32_t)lane, (uint32_t)RT)) - k;
        const uint32_t starts = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
        const uint32_t owner = min(owner0 + __popc(starts & (0xffffffffu >> (31 - lane))), (uint32_t)RT - 1u);
        // ---- 2. this lane's unit: rows 2u and 2u + 1 (if the bbox has it) of record `owner`, u = g - start
        const bool valid = (uint32_t)lane < ups && g < total;
        const uint32_t o16 = owner * 16u;
        const int4 c0 = vb200_lds128(aE1 + o16), c1 = vb200_lds128(aE2 + o16);
        const uint32_t unit = valid ? g - vb200_lds32(aStart + 4u * owner) : 0u;
        const int bx0 = c1.w & 0xff, w = (c1.w >> 16) & 0xff;
        const int y = ((c1.w >> 8) & 0xff) + 2 * (int)unit;
        const bool second = valid && 2u * unit + 1u < ((uint32_t)c1.w >> 24);
        // edge functions at the rows' last column, walked right to left: the sign bit pushed in at step i
        // ends up at bit (wmax - 1 - i); shifting the surplus of a narrower row out leaves column c at bit c
        const int xr = bx0 + w - 1;
        int e1 = c0.x * xr + c0.y * y + c0.z, e2 = c0.w * xr + c1.x * y + c1.y, e0 = c1.z - (e1 + e2);
        int f1 = e1 + c0.y, f2 = e2 + c1.x, f0 = c1.z - (f1 + f2);    // the row below
        const int A0 = -(c0.x + c0.w);
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, valid ? (uint32_t)w : 0u);
        uint32_t outA = 0, outB = 0;
#pragma unroll 2
        for(uint32_t i = 0; i < wmax; i++)
        {
          outA = __funnelshift_l((uint32_t)(e0 | e1 | e2), outA, 1);    // << 1 | (some edge function < 0)
          outB = __funnelshift_l((uint32_t)(f0 | f1 | f2), outB, 1);
          e0 -= A0;
          e1 -= c0.x;
          e2 -= c0.w;
          f0 -= A0;
          f1 -= c0.x;
          f2 -= c0.w;
        }
        // covered iff all three >= 0 (rasterizer.cpp:549), inside the row's w columns
        const uint32_t colmask = 0xffffffffu >> (32 - w);
        const uint32_t covA = valid ? (~(outA >> (wmax - (uint32_t)w)) & colmask) : 0u;
        const uint32_t covB = second ? (~(outB >> (wmax - (uint32_t)w)) & colmask) : 0u;
        uint32_t LA = __popc(covA), LB = __popc(covB);
        const int firstA = __ffs((int)covA) - 1, firstB = __ffs((int)covB) - 1;
        covered += LA + LB;
        // The three half-planes cut an interval out of a row, so its covered pixels are one run: first ..
        // first + L - 1. (Only wrapped int32 arithmetic on out-of-domain geometry could break that; such a row
        // is walked bit by bit instead, see below.)
        const bool raggedA = LA != 0u && (covA >> firstA) != (0xffffffffu >> (32 - LA));
        const bool raggedB = LB != 0u && (covB >> firstB) != (0xffffffffu >> (32 - LB));
        if(raggedA)
          LA = 0;
        if(raggedB)
          LB = 0;
        // ---- 3. the runs of the warp's rows end to end; rows without coverage drop out
        const uint32_t nzA = __ballot_sync(0xffffffffu, LA != 0u), nzB = __ballot_sync(0xffffffffu, LB != 0u);
        uint32_t end = LA + LB;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
          const uint32_t v = __shfl_up_sync(0xffffffffu, end, o);
          if(lane >= o)
            end += v;
        }
        const uint32_t hits = __shfl_sync(0xffffffffu, end, 31);
        const uint32_t nrec = __popc(nzA) + __popc(nzB);
        const uint32_t below = (1u << lane) - 1u;
        const uint32_t rank = __popc(nzA & below) + __popc(nzB & below);
        __syncwarp();    // the previous step's lookups in s_run are done
        // a run = {its first hit, record slot | tile pixel index of its first pixel << 8}
        if(LA != 0u)
          s_run[warp][rank] = make_uint2(end - LA - LB, owner | ((uint32_t)(y * VB200_TILE + bx0 + firstA) << 8));
        if(LB != 0u)
          s_run[warp][rank + (LA != 0u)] =
              make_uint2(end - LB, owner | ((uint32_t)((y + 1) * VB200_TILE + bx0 + firstB) << 8));
        __syncwarp();
        // run r starts at hit s_run[r].x; lane l keeps the starts of runs l and l + 32 for the owner lookups
        const uint32_t runStart0 = (uint32_t)lane < nrec ? s_run[warp][lane].x : 0xffffffffu;
        const uint32_t runStart1 = (uint32_t)lane + 32u < nrec ? s_run[warp][lane + 32].x : 0xffffffffu;
        // Hit h of a step that starts at hit hb belongs to run ownerBase + (number of runs that begin in
        // (hb, h]), where ownerBase is the run that owns hb itself: run 0 for the first step (it starts at hit
        // 0 and every run is non-empty), and for the next step the current one plus every run that begins in
        // (hb, hb + 32]. `begins` has bit r - 1 set when a run begins at hb + r.
        uint32_t ownerBase = 0;
        for(uint32_t hb = 0; hb < hits; hb += 32u)
        {
          const uint32_t rel0 = runStart0 - hb - 1u, rel1 = runStart1 - hb - 1u;
          const uint32_t begins =
              __reduce_or_sync(0xffffffffu, (rel0 < 32u ? (1u << rel0) : 0u) | (rel1 < 32u ? (1u << rel1) : 0u));
          const uint32_t r = min(ownerBase + __popc(begins & below), 63u);
          ownerBase += __popc(begins);
          const uint2 run = s_run[warp][r];
          const uint32_t h = hb + lane;
          const uint32_t slot16 = (run.y & (RT - 1u)) * 16u;
          const int4 t0 = vb200_lds128(aE1 + slot16), t1 = vb200_lds128(aE2 + slot16);
          const int idx = (int)((run.y >> 8) + (h - run.x)) & (VB200_TILE * VB200_TILE - 1);
          const int px = idx & 31, py = idx >> 5;
          const int b1 = t0.x * px + t0.y * py + t0.z;
          const int b2 = t0.w * px + t1.x * py + t1.y;
          const int b0 = t1.z - (b1 + b2);
          // the slot's current key is fetched before the depth arithmetic that decides whether it is needed
          const uint32_t aSlot = aVis + 8u * vb200_vis_slot((uint32_t)px, (uint32_t)py);
          const unsigned long long seen = vb200_lds64(aSlot);
          if(h >= hits)
            continue;
          float ia = 0.0f, z0 = 0.0f, z1 = 0.0f, z2 = 0.0f;
          if(MODE != VB200_RES_LAST_WINS || depthTest)
          {
            const int4 c2 = vb200_lds128(aZ + slot16);
            ia = __int_as_float(c2.x);
            z0 = __int_as_float(c2.y);
            z1 = __int_as_float(c2.z);
            z2 = __int_as_float(c2.w);
          }
          unsigned long long key;
          if(vb200_fragment_key<MODE>(b0, b1, b2, ia, z0, z1, z2, vb200_lds32(aKey + (slot16 >> 2)), depthTest,
                                      rs.depth_op, s_depth, idx, key))
            vb200_vis_min(aSlot, seen, key);
        }
        if(__any_sync(0xffffffffu, raggedA || raggedB))
        {
          // out-of-domain geometry only: a row with holes; its set bits are visited one by one
          const int4 c2 = vb200_lds128(aZ + o16);
          const uint32_t id = vb200_lds32(aKey + 4u * owner);
          for(int rowSel = 0; rowSel < 2; rowSel++)
          {
            uint32_t bits = rowSel ? (raggedB ? covB : 0u) : (raggedA ? covA : 0u);
            const int py = y + rowSel;
            while(bits)
            {
              const int px = bx0 + __ffs((int)bits) - 1;
              bits &= bits - 1u;
              const int b1 = c0.x * px + c0.y * py + c0.z, b2 = c0.w * px + c1.x * py + c1.y, b0 = c1.z - (b1 + b2);
              const int idx = (py * VB200_TILE + px) & (VB200_TILE * VB200_TILE - 1);
              unsigned long long key;
              if(vb200_fragment_key<MODE>(b0, b1, b2, __int_as_float(c2.x), __int_as_float(c2.y),
                                          __int_as_float(c2.z), __int_as_float(c2.w), id, depthTest, rs.depth_op,
                                          s_depth, idx, key))
              {
                const uint32_t aSlot = aVis + 8u * vb200_vis_slot((uint32_t)idx & 31u, (uint32_t)idx >> 5);
                vb200_vis_min(aSlot, vb200_lds64(aSlot), key);
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- phase B: shade the winner of every pixel, write back. Thread (warp, lane) owns pixel column `lane`
  // of rows warp, warp + RW, warp + 2 RW, ...
  const bool remote = p.mc_color != nullptr || p.num_peers != 0u;
  uint32_t aPw = vb200_smem_addr(s_pw), aSv = vb200_smem_addr(s_sv);
  asm volatile("" : "+r"(aPw), "+r"(aSv));
  const bool xin = tileX0 + lane < (int)rs.width;
  const uint32_t rowStep = (uint32_t)RW * rs.width;    // pixel offsets fit 32 bits (targets are at most 8192 x 8192)
  // `record(id, ly, y)`: the winner's barycentric numerators at the pixel and its per-triangle constants
  auto shade_rows = [&](auto record) {
    uint32_t gi = (uint32_t)(tileY0 + warp) * rs.width + (uint32_t)(tileX0 + lane);
    VB200_UNROLL(VB200_PB_UNROLL)
    for(int j = 0; j < VB200_TILE / RW; j++, gi += rowStep)
    {
      const int ly = warp + RW * j;
      const unsigned long long key = vb200_lds64(aVis + 8u * vb200_vis_slot(lane, ly));
      const uint32_t low = (uint32_t)key;
      bool won;
      uint32_t id;
      if(MODE == VB200_RES_MIN_FIRST || MODE == VB200_RES_MAX_FIRST)
      {
        won = low != 0u;
        id = low;
      }
      else if(MODE == VB200_RES_LAST_WINS)
      {
        won = key != ~0ull;
        id = ~low;
      }
      else
      {
        won = low != 0xffffffffu;
        id = ~low;
      }
      if(!xin || tileY0 + ly >= (int)rs.height)
        continue;
      if(!won)
      {
        if(clearColor)
          vb200_store_color(p, gi, p.clear_color, remote);
        if(clearDepth && rs.has_depth)
          __stcs(p.depth + gi, p.clear_depth);
        continue;
      }
      shaded++;
      int b0, b1, b2;
      float invarea, d0, d1, d2, invw0, invw1, invw2;
      const float4 *v0, *v1, *v2;
      record(id, ly, b0, b1, b2, invarea, d0, d1, d2, invw0, invw1, invw2, v0, v1, v2);
      // rasterizer.cpp:552-558, 581-588
      float n0 = __fmul_rn((float)b0, invarea);
      float n1 = __fmul_rn((float)b1, invarea);
      float n2 = __fmul_rn((float)b2, invarea);
      const float pixdepth = __fadd_rn(__fadd_rn(__fmul_rn(n0, d0), __fmul_rn(n1, d1)), __fmul_rn(n2, d2));
      n0 = __fmul_rn(n0, invw0);
      n1 = __fmul_rn(n1, invw1);
      n2 = __fmul_rn(n2, invw2);
      // 1.0f / x, correctly rounded (the dedicated reciprocal is a shorter sequence than the general division
      // and returns the same bits: both are the IEEE-rounded quotient)
      const float invlen = __frcp_rn(__fadd_rn(__fadd_rn(n0, n1), n2));
      n0 = __fmul_rn(n0, invlen);
      n1 = __fmul_rn(n1, invlen);
      n2 = __fmul_rn(n2, invlen);
      const float4 pix = vb200_fs(&env, n0, n1, n2, v0, v1, v2).color;    // (no OpKill here: runtime.cpp sends those shaders to the ordered kernel)
      vb200_store_color(p, gi, vb200_blend_store(rs, pix, clearColor ? p.clear_color : p.color[gi]), remote);
      if(depthWrite)
        __stcs(p.depth + gi, pixdepth);
      else if(clearDepth && rs.has_depth)
        __stcs(p.depth + gi, p.clear_depth);
    }
  };
  // the winner's edge values at this pixel from its staged record (int32 ring arithmetic, so this and the
  // reference's formulation, rasterizer.cpp:303-309,545-558, give the same bits)
  auto staged_record = [&](uint32_t slot16, int ly, int &b0, int &b1, int &b2, float &invarea, float &d0, float &d1,
                           float &d2, float &invw0, float &invw1, float &invw2, uint32_t &s0) {
    const int4 c0 = vb200_lds128(aE1 + slot16), c1 = vb200_lds128(aE2 + slot16), c2 = vb200_lds128(aZ + slot16),
               c4 = vb200_lds128(aPw + slot16);
    b1 = c0.x * lane + c0.y * ly + c0.z;
    b2 = c0.w * lane + c1.x * ly + c1.y;
    b0 = c1.z - (b1 + b2);
    invarea = __int_as_float(c2.x); d0 = __int_as_float(c2.y); d1 = __int_as_float(c2.z); d2 = __int_as_float(c2.w);
    invw0 = __int_as_float(c4.x); invw1 = __int_as_float(c4.y); invw2 = __int_as_float(c4.z);
    s0 = (uint32_t)c4.w;
  };
  if(recordsInSmem)
    shade_rows([&](uint32_t id, int ly, int &b0, int &b1, int &b2, float &invarea, float &d0, float &d1, float &d2,
                   float &invw0, float &invw1, float &invw2, const float4 *&v0, const float4 *&v1, const float4 *&v2) {
      const uint32_t slot16 = (id & (RT - 1u)) * 16u;
      uint32_t s0;
      staged_record(slot16, ly, b0, b1, b2, invarea, d0, d1, d2, invw0, invw1, invw2, s0);
      const unsigned long long c5 = vb200_lds64(aSv + (slot16 >> 1));
      v0 = p.interps + (size_t)s0 * rs.nslots;
      v1 = p.interps + (size_t)(uint32_t)c5 * rs.nslots;
      v2 = p.interps + (size_t)(uint32_t)(c5 >> 32) * rs.nslots;
    });
  else
    shade_rows([&](uint32_t id, int ly, int &b0, int &b1, int &b2, float &invarea, float &d0, float &d1, float &d2,
                   float &invw0, float &invw1, float &invw2, const float4 *&v0, const float4 *&v1, const float4 *&v2) {
      // several rounds: the record is gathered again, edge values exactly as rasterizer.cpp:303-309,545-558
      const Vb200TriSetup su = vb200_load_setup(p, (rs.slot_keys ? (id >> 8) : id) - 1u);
      const int ABx = su.x1 - su.x0, ABy = su.y1 - su.y0, ACx = su.x2 - su.x0, ACy = su.y2 - su.y0;
      const int area2 = ABx * ACy - ABy * ACx;
      const int sgn = area2 > 0 ? 1 : -1;
      const int PAx = su.x0 - (tileX0 + lane), PAy = su.y0 - (tileY0 + ly);
      const int ux = ACx * PAy - ACy * PAx, uy = PAx * ABy - PAy * ABx;
      b0 = (area2 - (ux + uy)) * sgn; b1 = ux * sgn; b2 = uy * sgn;
      invarea = su.invarea; d0 = su.d0; d1 = su.d1; d2 = su.d2;
      invw0 = su.invw0; invw1 = su.invw1; invw2 = su.invw2;
      v0 = p.interps + (size_t)su.s0 * rs.nslots;
      v1 = p.interps + (size_t)su.s1 * rs.nslots;
      v2 = p.interps + (size_t)su.s2 * rs.nslots;
    });
  if(rs.count_fragments)
    vb200_count_fragments(p.counters, covered, shaded);
}

#define VB200_RESOLVE_KERNEL(NAME, MODE)                                                              \
  extern "C" __global__ void __launch_bounds__(VB200_RESOLVE_THREADS)                                \
      NAME(const __grid_constant__ Vb200Env env, const __grid_constant__ Vb200TileParams p)          \
  {                                                                                                   \
    vb200_tile_resolve_body<MODE>(env, p);                                                            \
  }
VB200_RESOLVE_KERNEL(vb200_k_tile_resolve_min_first, VB200_RES_MIN_FIRST)
VB200_RESOLVE_KERNEL(vb200_k_tile_resolve_min_last, VB200_RES_MIN_LAST)
VB200_RESOLVE_KERNEL(vb200_k_tile_resolve_max_first, VB200_RES_MAX_FIRST)
VB200_RESOLVE_KERNEL(vb200_k_tile_resolve_max_last, VB200_RES_MAX_LAST)
VB200_RESOLVE_KERNEL(vb200_k_tile_resolve_last_wins, VB200_RES_LAST_WINS)
