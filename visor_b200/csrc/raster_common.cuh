// raster_common.cuh — device helpers shared by scaffold.cu (JIT-linked kernels) and fixed.cu.
// All float arithmetic is explicit round-to-nearest, unfused, in the reference's order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "device_types.h"
#include "kernels.h"

// int(float) as the reference's x86-64 build performs it (cvttss2si): truncation toward zero,
// 0x80000000 for NaN and out-of-range inputs.
__device__ __forceinline__ int vb200_cvtt(float f)
{
  return (fabsf(f) < 2147483648.0f) ? __float2int_rz(f) : (int)0x80000000;
}

__device__ __forceinline__ float vb200_ld_f32_unaligned(const uint8_t *p)
{
  if((((uintptr_t)p) & 3) == 0)
    return __ldg((const float *)p);
  uint32_t u = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) |
               ((uint32_t)__ldg(p + 3) << 24);
  return __uint_as_float(u);
}

// a triangle as the tile kernels use it, from its 16-byte record q and the three raster vertices a, b, c
__device__ __forceinline__ Vb200TriSetup vb200_unpack_setup(const int4 &q, const int4 &a, const int4 &b, const int4 &c)
{
  Vb200TriSetup r;
  r.x0 = a.x; r.y0 = a.y; r.x1 = b.x; r.y1 = b.y; r.x2 = c.x; r.y2 = c.y;
  r.invw0 = __int_as_float(a.z); r.invw1 = __int_as_float(b.z); r.invw2 = __int_as_float(c.z);
  r.d0 = __int_as_float(a.w); r.d1 = __int_as_float(b.w); r.d2 = __int_as_float(c.w);
  r.s0 = (uint32_t)q.x; r.s1 = (uint32_t)q.y; r.s2 = (uint32_t)q.z; r.invarea = __int_as_float(q.w);
  return r;
}

__device__ __forceinline__ Vb200TriSetup vb200_load_setup(const Vb200TileParams &p, uint32_t t)
{
  // 16-byte triangle record, then one 16-byte read-only gather per corner
  const int4 q = __ldg((const int4 *)(p.tri + t));
  const int4 a = __ldg((const int4 *)(p.rv + (uint32_t)q.x));
  const int4 b = __ldg((const int4 *)(p.rv + (uint32_t)q.y));
  const int4 c = __ldg((const int4 *)(p.rv + (uint32_t)q.z));
  Vb200TriSetup r;
  r.x0 = a.x; r.y0 = a.y; r.x1 = b.x; r.y1 = b.y; r.x2 = c.x; r.y2 = c.y;
  r.invw0 = __int_as_float(a.z); r.invw1 = __int_as_float(b.z); r.invw2 = __int_as_float(c.z);
  r.d0 = __int_as_float(a.w); r.d1 = __int_as_float(b.w); r.d2 = __int_as_float(c.w);
  r.s0 = (uint32_t)q.x; r.s1 = (uint32_t)q.y; r.s2 = (uint32_t)q.z; r.invarea = __int_as_float(q.w);
  return r;
}

// does the packed inclusive tile range (tx0 | ty0 << 8 | tx1 << 16 | ty1 << 24) contain tile (tx, ty)?
// VB200_TILES_DEAD has tx0 = 255 > tx1 = 0 and contains nothing.
__device__ __forceinline__ bool vb200_tile_in_range(uint32_t tiles, uint32_t tx, uint32_t ty)
{
  return tx >= (tiles & 0xffu) && tx <= ((tiles >> 16) & 0xffu) && ty >= ((tiles >> 8) & 0xffu) && ty <= (tiles >> 24);
}

// rasterizer.cpp:562-576
__device__ __forceinline__ bool vb200_depth_pass(uint32_t op, float pixdepth, float curdepth)
{
  switch(op)
  {
    case 0: return false;                   // NEVER
    case 1: return pixdepth < curdepth;     // LESS
    case 2: return pixdepth == curdepth;    // EQUAL
    case 3: return pixdepth <= curdepth;    // LESS_OR_EQUAL
    case 4: return pixdepth > curdepth;     // GREATER
    case 5: return pixdepth != curdepth;    // NOT_EQUAL
    case 6: return pixdepth >= curdepth;    // GREATER_OR_EQUAL
    default: return true;
  }
}

__device__ __forceinline__ float vb200_clamp01(float in)
{
  return in > 1.0f ? 1.0f : (in < 0.0f ? 0.0f : in);    // rasterizer.cpp:277-280
}

__device__ __forceinline__ float vb200_blend_factor(uint32_t f, float alpha)
{
  switch(f)    // rasterizer.cpp:601-653: four factors, anything else stays 1.0
  {
    case 0: return 0.0f;                      // ZERO
    case 1: return 1.0f;                      // ONE
    case 6: return alpha;                     // SRC_ALPHA
    case 7: return __fsub_rn(1.0f, alpha);    // ONE_MINUS_SRC_ALPHA
    default: return 1.0f;
  }
}

// blend (rasterizer.cpp:593-672) + truncating BGR store, alpha byte untouched (:674-676).
// `cur` is the pixel's 32-bit word: byte0 = B, byte1 = G, byte2 = R, byte3 = A.
__device__ __forceinline__ uint32_t vb200_blend_store(const Vb200RasterState &rs, float4 pix, uint32_t cur)
{
  if(rs.blend_enable)
  {
    const float ex = __fdiv_rn((float)((cur >> 16) & 0xffu), 255.0f);
    const float ey = __fdiv_rn((float)((cur >> 8) & 0xffu), 255.0f);
    const float ez = __fdiv_rn((float)(cur & 0xffu), 255.0f);
    const float srcF = vb200_blend_factor(rs.src_factor, pix.w);
    const float dstF = vb200_blend_factor(rs.dst_factor, pix.w);
    if(rs.blend_op == 0u)    // VK_BLEND_OP_ADD; other ops are outside the reference's defined domain
    {
      pix.x = __fadd_rn(__fmul_rn(srcF, pix.x), __fmul_rn(dstF, ex));
      pix.y = __fadd_rn(__fmul_rn(srcF, pix.y), __fmul_rn(dstF, ey));
      pix.z = __fadd_rn(__fmul_rn(srcF, pix.z), __fmul_rn(dstF, ez));
    }
  }
  // byte(clamp01(c) * 255.0f): __saturatef clamps to [0,1] in one instruction; it maps NaN to 0 where
  // clamp01 keeps NaN, but byte(NaN * 255) is 0 on the reference's x86 too, so the stored byte is equal
  const uint32_t r = (uint32_t)__float2int_rz(__fmul_rn(__saturatef(pix.x), 255.0f));
  const uint32_t g = (uint32_t)__float2int_rz(__fmul_rn(__saturatef(pix.y), 255.0f));
  const uint32_t b = (uint32_t)__float2int_rz(__fmul_rn(__saturatef(pix.z), 255.0f));
  return (cur & 0xff000000u) | (r << 16) | (g << 8) | b;
}

// colour store of the tile kernels: local image + every other rank's image (fused sort-first exchange),
// through one switch-replicated multicast store when an NVLS mapping is set, else one store per peer.
// `remote` (uniform: some exchange target is set) keeps the single-GPU path at one store and one branch.
__device__ __forceinline__ void vb200_store_color(const Vb200TileParams &p, uint32_t gi, uint32_t v, bool remote)
{
  __stcs(p.color + gi, v);    // streaming store: the frame must not sweep the geometry out of L2
  if(!remote)
    return;
  if(p.mc_color)
    asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(p.mc_color + gi), "r"(v) : "memory");
  else
  {
    // (not unrolled: this function is inlined at every store of the tile kernels, and ptxas turned the
    // seven-peer loop into ~170 instructions per site, a third of the resolve kernels' code)
#pragma unroll 1
    for(uint32_t r = 0; r < p.num_peers; r++)
      p.peer_color[r][gi] = v;
  }
}

__device__ __forceinline__ void vb200_count_fragments(Vb200DrawCounters *c, uint32_t covered, uint32_t shaded)
{
  for(int o = 16; o > 0; o >>= 1)
  {
    covered += __shfl_down_sync(0xffffffffu, covered, o);
    shaded += __shfl_down_sync(0xffffffffu, shaded, o);
  }
  if((threadIdx.x & 31) == 0 && (covered | shaded))
  {
    Vb200DrawCounters::Slot &s = c->slot[blockIdx.x & (VB200_COUNTER_SLOTS - 1)];
    atomicAdd(&s.fragments_covered, (unsigned long long)covered);
    atomicAdd(&s.fragments_shaded, (unsigned long long)shaded);
  }
}

// ---- texture unit --------------------------------------------------------------------------
// 5:6:5 endpoint expansion of the BC decoder the reference uses (3rdparty/decompress.c:122-134):
// ((t/32 + t)/32) with t = c5*255 + 16, ((t/64 + t)/64) with t = c6*255 + 32; packed as 0x00BBGGRR
__device__ __forceinline__ uint32_t vb200_bc_endpoint(uint32_t c)
{
  uint32_t t = (c >> 11) * 255u + 16u;
  const uint32_t r = ((t >> 5) + t) >> 5;
  t = ((c >> 5) & 0x3fu) * 255u + 32u;
  const uint32_t g = ((t >> 6) + t) >> 6;
  t = (c & 0x1fu) * 255u + 16u;
  const uint32_t b = ((t >> 5) + t) >> 5;
  return (r & 0xffu) | ((g & 0xffu) << 8) | ((b & 0xffu) << 16);
}

// One texel of a BC2/BC3 block (16 B: 8 B alpha + 8 B colour), as DecompressBlockBC2/BC3 produce it
// (decompress.c:233-350). BC2 goes through the BC1 colour path, which keeps BC1's three-colour mode
// when color0 <= color1; BC3 always interpolates four colours. Returns packed RGBA bytes.
__device__ __forceinline__ uint32_t vb200_bc_texel(const uint8_t *blk, bool bc3, int x, int y)
{
  uint4 w;
  if((((uintptr_t)blk) & 15) == 0)
    w = __ldg((const uint4 *)blk);
  else    // images may be bound at any byte offset (alignment 1, images.cpp:51-56)
    w = make_uint4(__float_as_uint(vb200_ld_f32_unaligned(blk)), __float_as_uint(vb200_ld_f32_unaligned(blk + 4)),
                   __float_as_uint(vb200_ld_f32_unaligned(blk + 8)), __float_as_uint(vb200_ld_f32_unaligned(blk + 12)));
  const int ti = (y & 3) * 4 + (x & 3);
  uint32_t alpha;
  if(!bc3)
  {
    const uint32_t row = ((y & 2) ? w.y : w.x) >> ((y & 1) * 16);
    alpha = ((row >> (4 * (x & 3))) & 0xfu) * 17u;
  }
  else
  {
    const uint32_t alpha0 = w.x & 0xffu, alpha1 = (w.x >> 8) & 0xffu;
    // 48 bits of 3-bit codes start at byte 2: two 24-bit groups of eight codes (decompress.c:89-107)
    const unsigned long long bits = (((unsigned long long)w.y << 32) | w.x) >> 16;
    const uint32_t ac = (uint32_t)(bits >> (3 * ti)) & 7u;
    if(ac == 0u)
      alpha = alpha0;
    else if(ac == 1u)
      alpha = alpha1;
    else if(alpha0 > alpha1)
      alpha = ((8u - ac) * alpha0 + (ac - 1u) * alpha1) / 7u;
    else if(ac == 6u)
      alpha = 0u;
    else if(ac == 7u)
      alpha = 255u;
    else
      alpha = ((6u - ac) * alpha0 + (ac - 1u) * alpha1) / 5u;
  }
  const uint32_t color0 = w.z & 0xffffu, color1 = w.z >> 16;
  const uint32_t e0 = vb200_bc_endpoint(color0), e1 = vb200_bc_endpoint(color1);
  const uint32_t pc = (w.w >> (2 * ti)) & 3u;
  uint32_t rgb = 0;
#pragma unroll
  for(int c = 0; c < 3; c++)
  {
    const uint32_t a = (e0 >> (8 * c)) & 0xffu, b = (e1 >> (8 * c)) & 0xffu;
    uint32_t v;
    if(bc3 || color0 > color1)
      v = pc == 0u ? a : pc == 1u ? b : pc == 2u ? (2u * a + b) / 3u : (a + 2u * b) / 3u;
    else
      v = pc == 0u ? a : pc == 1u ? b : pc == 2u ? (a + b) / 2u : 0u;
    rgb |= (v & 0xffu) << (8 * c);
  }
  return rgb | ((alpha & 0xffu) << 24);
}

// One texel as CacheCoord converts it (texture_sampling.cpp:92-133): channel c = float(byte[c])/255.0f.
// Linear formats: address base + (y*width + x)*bpp. BC2/BC3 (VkFormat 135/137): block (x>>2, y>>2) of
// a (width>>2)-block-wide image, decoded per texel. The reference's 4x4 LRU block cache is a pure
// cache; here the read-only L1/texture path (ld.global.nc) plays that role.
// float(byte) / 255.0f: an IEEE division per channel in the reference. `lut` (256 floats holding exactly
// those quotients, in shared memory) replaces the sixteen divisions of a bilinear sample by loads.
#ifndef VB200_UNORM_NEWTON
#define VB200_UNORM_NEWTON 1    // 0: look the quotient up in the shared-memory table instead (measured: C5 +6 %, C2 +2 %)
#endif
__device__ __forceinline__ float vb200_unorm8(const float *lut, uint32_t b)
{
#if VB200_UNORM_NEWTON
  // float(b) / 255.0f without a division or a table. 1/255 = r + e with r = RN(1/255) and e = RN(1/255 - r):
  // b*r is exact inside the fma and b*e is far below its last bit, so fma(b, r, b*e) is the correctly rounded
  // quotient for every byte value (tests/test_gpu_parity.py test_unorm8_conversion_is_the_ieee_quotient compares
  // all 256 with the IEEE division on the GPU, tests/test_abi.py checks the constants with exact rationals).
  // Two operations per channel; a 256-entry table of the quotients in shared memory costs a 3-4 way bank
  // conflict per lookup on random texels, sixteen lookups per bilinear sample: 47 M conflicts per C5 frame.
  const float f = (float)b;
  const float r = 0.0039215688593685626983642578125f;       // RN(1/255) = 0x1.010102p-8
  const float e = -2.31917582360630104257e-10f;              // RN(1/255 - r) = -0x1.fdfdfep-33
  return __fmaf_rn(f, r, __fmul_rn(f, e));
#else
  return lut ? lut[b] : __fdiv_rn((float)b, 255.0f);
#endif
}

__device__ __forceinline__ float4 vb200_unorm8x4(const float *lut, uint32_t u)
{
  return make_float4(vb200_unorm8(lut, u & 0xffu), vb200_unorm8(lut, (u >> 8) & 0xffu),
                     vb200_unorm8(lut, (u >> 16) & 0xffu), vb200_unorm8(lut, u >> 24));
}

__device__ __forceinline__ float4 vb200_texel(const uint8_t *base, uint32_t width, uint32_t bpp, uint32_t format,
                                              int x, int y, const float *lut)
{
  uint32_t u;
  if(format == 135u || format == 137u)
    u = vb200_bc_texel(base + ((size_t)(y >> 2) * (width >> 2) + (size_t)(x >> 2)) * 16, format == 137u, x, y);
  else
  {
    const uint8_t *p = base + ((size_t)y * width + (size_t)x) * bpp;
    if(bpp == 4)
      u = __ldg((const uint32_t *)p);
    else
      u = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) |
          ((uint32_t)__ldg(p + 3) << 24);
  }
  return vb200_unorm8x4(lut, u);
}

// sample_tex_wrapped (texture_sampling.cpp:139-184): repeat wrap, bilinear, mip 0, no half-texel offset
__device__ __forceinline__ float4 vb200_sample_tex_impl(float u, float v, const Vb200Image *img,
                                                        unsigned long long byteOffs, const float *lut = nullptr)
{
  const uint32_t width = img->width, height = img->height, bpp = img->bpp, fmt = img->format;
  const uint8_t *base = img->pixels + byteOffs;
  u = __fsub_rn(u, floorf(u));
  v = __fsub_rn(v, floorf(v));
  u = __fmul_rn(u, (float)width);
  v = __fmul_rn(v, (float)height);
  const int iu0 = vb200_cvtt(u), iv0 = vb200_cvtt(v);
  int iu1 = iu0 + 1, iv1 = iv0 + 1;
  if(iu1 >= (int)width)
    iu1 -= (int)width;
  if(iv1 >= (int)height)
    iv1 -= (int)height;
  const float fu = __fsub_rn(u, (float)iu0), fv = __fsub_rn(v, (float)iv0);
  const float inv_fu = __fsub_rn(1.0f, fu), inv_fv = __fsub_rn(1.0f, fv);
  float4 TL, TR, BL, BR;
  if(bpp == 4u && fmt != 135u && fmt != 137u && (((uintptr_t)base) & 3) == 0)
  {
    // the common case decided once per sample instead of once per texel: linear 4-byte texels at an aligned
    // base. Texel indices fit 32 bits (images are at most 8192 x 8192), one 64-bit address per row.
    const uint32_t *row0 = (const uint32_t *)base + (uint32_t)iv0 * width;
    const uint32_t *row1 = (const uint32_t *)base + (uint32_t)iv1 * width;
    const uint32_t tl = __ldg(row0 + (uint32_t)iu0), tr = __ldg(row0 + (uint32_t)iu1);
    const uint32_t bl = __ldg(row1 + (uint32_t)iu0), br = __ldg(row1 + (uint32_t)iu1);
    TL = vb200_unorm8x4(lut, tl);
    TR = vb200_unorm8x4(lut, tr);
    BL = vb200_unorm8x4(lut, bl);
    BR = vb200_unorm8x4(lut, br);
  }
  else
  {
    TL = vb200_texel(base, width, bpp, fmt, iu0, iv0, lut);
    TR = vb200_texel(base, width, bpp, fmt, iu1, iv0, lut);
    BL = vb200_texel(base, width, bpp, fmt, iu0, iv1, lut);
    BR = vb200_texel(base, width, bpp, fmt, iu1, iv1, lut);
  }
  float4 top, bottom, out;
  top.x = __fadd_rn(__fmul_rn(TL.x, inv_fu), __fmul_rn(TR.x, fu));
  top.y = __fadd_rn(__fmul_rn(TL.y, inv_fu), __fmul_rn(TR.y, fu));
  top.z = __fadd_rn(__fmul_rn(TL.z, inv_fu), __fmul_rn(TR.z, fu));
  top.w = __fadd_rn(__fmul_rn(TL.w, inv_fu), __fmul_rn(TR.w, fu));
  bottom.x = __fadd_rn(__fmul_rn(BL.x, inv_fu), __fmul_rn(BR.x, fu));
  bottom.y = __fadd_rn(__fmul_rn(BL.y, inv_fu), __fmul_rn(BR.y, fu));
  bottom.z = __fadd_rn(__fmul_rn(BL.z, inv_fu), __fmul_rn(BR.z, fu));
  bottom.w = __fadd_rn(__fmul_rn(BL.w, inv_fu), __fmul_rn(BR.w, fu));
  out.x = __fadd_rn(__fmul_rn(top.x, inv_fv), __fmul_rn(bottom.x, fv));
  out.y = __fadd_rn(__fmul_rn(top.y, inv_fv), __fmul_rn(bottom.y, fv));
  out.z = __fadd_rn(__fmul_rn(top.z, inv_fv), __fmul_rn(bottom.z, fv));
  out.w = __fadd_rn(__fmul_rn(top.w, inv_fv), __fmul_rn(bottom.w, fv));
  return out;
}

// sample_cube_wrapped (texture_sampling.cpp:186-250): six tests in sequence, later matches win;
// layer offset = CalcSubresourceByteOffset(tex, 0, face) = face * full-mip-chain size.
__device__ __forceinline__ float4 vb200_sample_cube_impl(float x, float y, float z, const Vb200Image *img,
                                                         const float *lut = nullptr)
{
  const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
  const bool px = x > 0.0f, py = y > 0.0f, pz = z > 0.0f;
  float axis = 0.0f, u = 0.0f, v = 0.0f;
  uint32_t face = 0;
  if(px && ax >= ay && ax >= az)  { axis = ax; u = -z; v = -y; face = 0; }
  if(!px && ax >= ay && ax >= az) { axis = ax; u = z;  v = -y; face = 1; }
  if(py && ay >= ax && ay >= az)  { axis = ay; u = x;  v = z;  face = 2; }
  if(!py && ay >= ax && ay >= az) { axis = ay; u = x;  v = -z; face = 3; }
  if(pz && az >= ax && az >= ay)  { axis = az; u = x;  v = -y; face = 4; }
  if(!pz && az >= ax && az >= ay) { axis = az; u = -x; v = -y; face = 5; }
  const float su = __fmul_rn(0.5f, __fadd_rn(__fdiv_rn(u, axis), 1.0f));
  const float sv = __fmul_rn(0.5f, __fadd_rn(__fdiv_rn(v, axis), 1.0f));
  return vb200_sample_tex_impl(su, sv, img, (unsigned long long)face * img->slice_bytes, lut);
}
