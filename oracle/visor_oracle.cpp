/*
 * oracle/visor_oracle.cpp — TEST INFRASTRUCTURE ONLY (parity oracle; never on the product path).
 *
 * CPU restatement of visor's draw-execution path in its deterministic (serial) order:
 *   ClearTarget            rasterizer.cpp:312-361
 *   GetIndex / ShadeVerts  rasterizer.cpp:100-237   (triangle list + strip assembly)
 *   ToWindow               rasterizer.cpp:239-253
 *   DrawTriangles setup    rasterizer.cpp:385-452   (area, cull, bbox clamp, invarea, invw, depth)
 *   ProcessTriangles       rasterizer.cpp:522-696   (coverage, depth test, perspective, FS, blend, store)
 *   sample_tex_wrapped     texture_sampling.cpp:139-184 (+ texel fetch of CacheCoord :119-134)
 *   sample_cube_wrapped    texture_sampling.cpp:186-250
 *   CalcSubresourceByteOffset precompiled.cpp:3-36
 * The shader stage is oracle/spirv_cpu.cpp.
 *
 * Order: the reference cuts each triangle's bbox into 32x32 blocks pushed to a FIFO which the main
 * thread drains after all triangles are queued (rasterizer.cpp:454-515, no worker threads = the
 * deterministic mode).  Blocks of one triangle are disjoint, so per pixel the
 * fragments arrive in triangle order; this file therefore walks each triangle's whole bbox in turn.
 * The 4x4 texel LRU (texture_sampling.cpp:5-90) is a pure cache and is not modelled.
 *
 * PARITY STATUS: pinned against the reference's own rasterizer.cpp / texture_sampling.cpp compiled
 * unmodified (oracle/_ref, see oracle/Makefile) on randomised scenes — tests/test_oracle_vs_ref.py —
 * and against the committed fixtures in tests/golden/ that were generated from that build.
 * The interface is the product's C-ABI structs (include/visor_b200.h) with a vor_ prefix.
 *
 * Build: g++ -O2 -ffp-contract=off (no -ffast-math, no -march=native): IEEE binary32, unfused.
 */
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../include/visor_b200.h"
#include "spirv_cpu.h"

namespace
{
typedef unsigned char byte;

enum
{
  TOPOLOGY_TRIANGLE_LIST = 3,
  TOPOLOGY_TRIANGLE_STRIP = 4,
  FRONT_FACE_CLOCKWISE = 1,
  CULL_FRONT = 1,
  CULL_BACK = 2,
  INDEX_UINT16 = 0,
  CMP_NEVER = 0, CMP_LESS = 1, CMP_EQUAL = 2, CMP_LEQUAL = 3, CMP_GREATER = 4, CMP_NOTEQUAL = 5,
  CMP_GEQUAL = 6, CMP_ALWAYS = 7,
  BF_ZERO = 0, BF_ONE = 1, BF_SRC_ALPHA = 6, BF_ONE_MINUS_SRC_ALPHA = 7,
  BLEND_OP_ADD = 0,
};

std::string g_err;

struct Vertex
{
  float position[4];
  float interps[10][4];
};

// ---- texture unit ------------------------------------------------------------------------

uint64_t subresourceOffset(const vb200_image *img, uint32_t mip, uint32_t layer)
{
  // precompiled.cpp:3-36
  uint64_t offs = 0;
  const uint32_t w = img->width, h = img->height, bpp = img->bytes_per_pixel;
  for(uint32_t m = 0; m < mip; m++)
  {
    uint32_t mw = w >> m, mh = h >> m;
    if(mw < 1)
      mw = 1;
    if(mh < 1)
      mh = 1;
    offs += mw * mh * bpp;
  }
  if(layer > 0)
  {
    uint32_t mw = w, mh = h;
    uint64_t slice = 0;
    for(uint32_t m = 0; m < img->mip_levels; m++)
    {
      slice += mw * mh * bpp;
      mw = mw >> 1 > 1 ? mw >> 1 : 1;
      mh = mh >> 1 > 1 ? mh >> 1 : 1;
    }
    offs += slice * layer;
  }
  return offs;
}

// ---- block-compressed texels ------------------------------------------------------------------
// Restates the decoder the reference calls for BC2/BC3 (3rdparty/decompress.c, "Anteru" BC decoder),
// for ONE texel of a 16-byte block instead of the whole 4x4 block.
// 5:6:5 endpoint expansion (decompress.c:122-134): r = ((t/32 + t)/32) with t = c5*255 + 16 etc.
void bcEndpoints(uint16_t c, uint32_t rgb[3])
{
  uint32_t t = (uint32_t)(c >> 11) * 255 + 16;
  rgb[0] = (uint8_t)((t / 32 + t) / 32);
  t = (uint32_t)((c & 0x07E0) >> 5) * 255 + 32;
  rgb[1] = (uint8_t)((t / 64 + t) / 64);
  t = (uint32_t)(c & 0x001F) * 255 + 16;
  rgb[2] = (uint8_t)((t / 32 + t) / 32);
}

// colour half of a block (8 bytes): decompress.c:111-198 when `threeColourMode` may apply (BC2 goes
// through DecompressBlockBC1Internal, which keeps BC1's color0 <= color1 mode), :246-316 for BC3
// (always four colours)
void bcColour(const byte *blk, int texelIdx, bool bc1Modes, byte rgb[3])
{
  const uint16_t color0 = (uint16_t)(blk[0] | (blk[1] << 8)), color1 = (uint16_t)(blk[2] | (blk[3] << 8));
  uint32_t e0[3], e1[3];
  bcEndpoints(color0, e0);
  bcEndpoints(color1, e1);
  const uint32_t code = (uint32_t)blk[4] | ((uint32_t)blk[5] << 8) | ((uint32_t)blk[6] << 16) | ((uint32_t)blk[7] << 24);
  const uint32_t pc = (code >> (2 * texelIdx)) & 3u;
  for(int c = 0; c < 3; c++)
  {
    uint32_t v;
    if(!bc1Modes || color0 > color1)
      v = pc == 0 ? e0[c] : pc == 1 ? e1[c] : pc == 2 ? (2 * e0[c] + e1[c]) / 3 : (e0[c] + 2 * e1[c]) / 3;
    else
      v = pc == 0 ? e0[c] : pc == 1 ? e1[c] : pc == 2 ? (e0[c] + e1[c]) / 2 : 0;
    rgb[c] = (byte)v;
  }
}

// one texel as the cache fill converts it (texture_sampling.cpp:92-133)
void texel(const vb200_image *tex, uint64_t byteOffs, int x, int y, float out[4])
{
  const byte *base = (const byte *)tex->pixels + byteOffs;
  if(tex->format == 135u || tex->format == 137u)    // VK_FORMAT_BC2_UNORM_BLOCK / BC3_UNORM_BLOCK
  {
    // texture_sampling.cpp:96-118: block (x>>2, y>>2) of a (width>>2)-block-wide image, 16 B per block
    const uint32_t widthInBlocks = tex->width >> 2;
    const byte *blk = base + ((uint64_t)(y >> 2) * widthInBlocks + (uint64_t)(x >> 2)) * 16;
    const int ti = (y & 3) * 4 + (x & 3);
    byte rgba[4];
    if(tex->format == 135u)
    {
      // DecompressBlockBC2 (decompress.c:330-350): 4-bit alpha * 17, then the BC1 colour block
      const uint16_t row = (uint16_t)(blk[2 * (y & 3)] | (blk[2 * (y & 3) + 1] << 8));
      rgba[3] = (byte)(((row >> (4 * (x & 3))) & 0xF) * 17);
      bcColour(blk + 8, ti, true, rgba);
    }
    else
    {
      // DecompressBlockBC3 (decompress.c:233-319): two endpoints + 16 3-bit codes (:89-107)
      const uint32_t alpha0 = blk[0], alpha1 = blk[1];
      const byte *packed = blk + 2 + 3 * (ti >> 3);
      const uint32_t tmp = (uint32_t)packed[0] | ((uint32_t)packed[1] << 8) | ((uint32_t)packed[2] << 16);
      const int ac = (int)((tmp >> (3 * (ti & 7))) & 7u);
      uint32_t a;
      if(ac == 0)
        a = alpha0;
      else if(ac == 1)
        a = alpha1;
      else if(alpha0 > alpha1)
        a = ((8 - ac) * alpha0 + (ac - 1) * alpha1) / 7;
      else if(ac == 6)
        a = 0;
      else if(ac == 7)
        a = 255;
      else
        a = ((6 - ac) * alpha0 + (ac - 1) * alpha1) / 5;
      rgba[3] = (byte)a;
      bcColour(blk + 8, ti, false, rgba);
    }
    for(int c = 0; c < 4; c++)
      out[c] = float(rgba[c]) / 255.0f;
    return;
  }
  const uint32_t bpp = tex->bytes_per_pixel;
  const byte *p = base + ((uint64_t)y * tex->width + x) * bpp;
  for(int c = 0; c < 4; c++)
    out[c] = float(p[c]) / 255.0f;
}

void sampleTex(float u, float v, const vb200_image *tex, uint64_t byteOffs, float out[4])
{
  // texture_sampling.cpp:142-183
  u = u - floorf(u);
  v = v - floorf(v);
  u *= tex->width;
  v *= tex->height;
  int iu0 = int(u), iv0 = int(v);
  int iu1 = iu0 + 1, iv1 = iv0 + 1;
  if(iu1 >= (int)tex->width)
    iu1 -= tex->width;
  if(iv1 >= (int)tex->height)
    iv1 -= tex->height;
  float fu = u - float(iu0), fv = v - float(iv0);
  float inv_fu = 1.0f - fu, inv_fv = 1.0f - fv;
  float TL[4], TR[4], BL[4], BR[4];
  texel(tex, byteOffs, iu0, iv0, TL);
  texel(tex, byteOffs, iu1, iv0, TR);
  texel(tex, byteOffs, iu0, iv1, BL);
  texel(tex, byteOffs, iu1, iv1, BR);
  for(int c = 0; c < 4; c++)
  {
    float top = TL[c] * inv_fu + TR[c] * fu;
    float bottom = BL[c] * inv_fu + BR[c] * fu;
    out[c] = top * inv_fv + bottom * fv;
  }
}

void sampleCube(float x, float y, float z, const vb200_image *tex, float out[4])
{
  // texture_sampling.cpp:189-249: the six tests run in sequence, later matches overwrite earlier
  float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
  bool px = x > 0.0f, py = y > 0.0f, pz = z > 0.0f;
  float axis = 0.0f, u = 0.0f, v = 0.0f;
  uint64_t offset = 0;
  if(px && ax >= ay && ax >= az)  { axis = ax; u = -z; v = -y; offset = subresourceOffset(tex, 0, 0); }
  if(!px && ax >= ay && ax >= az) { axis = ax; u = z;  v = -y; offset = subresourceOffset(tex, 0, 1); }
  if(py && ay >= ax && ay >= az)  { axis = ay; u = x;  v = z;  offset = subresourceOffset(tex, 0, 2); }
  if(!py && ay >= ax && ay >= az) { axis = ay; u = x;  v = -z; offset = subresourceOffset(tex, 0, 3); }
  if(pz && az >= ax && az >= ay)  { axis = az; u = x;  v = -y; offset = subresourceOffset(tex, 0, 4); }
  if(!pz && az >= ax && az >= ay) { axis = az; u = -x; v = -y; offset = subresourceOffset(tex, 0, 5); }
  sampleTex(0.5f * (u / axis + 1.0f), 0.5f * (v / axis + 1.0f), tex, offset, out);
}

// ---- shader environment (restates the accessors of spirv_compile.cpp:552-627) -------------

const vb200_binding *findBinding(const vb200_draw_state *s, uint32_t set, uint32_t bind)
{
  for(uint32_t i = 0; i < s->num_bindings; i++)
    if(s->bindings[i].set == set && s->bindings[i].binding == bind)
      return &s->bindings[i];
  return NULL;
}

void envVertexAttr(void *user, uint32_t vertexIndex, uint32_t attr, float out[4])
{
  const vb200_draw_state *s = (const vb200_draw_state *)user;
  const vb200_vertex_attr &a = s->pipeline->vattrs[attr & 15];
  const byte *ptr = (const byte *)s->vbs[a.vb & 3].buffer.bytes + s->vbs[a.vb & 3].offset;
  ptr += a.offset;
  ptr += (uint64_t)a.stride * vertexIndex;
  if(!vor::fetch_vertex_attr(a.format, ptr, out))
  {
    fprintf(stderr, "vor: Unhandled vertex attribute format %u\n", a.format);
    abort();
  }
}
const uint8_t *envBufferPtr(void *user, uint32_t set, uint32_t bind)
{
  const vb200_binding *b = findBinding((const vb200_draw_state *)user, set, bind);
  return b ? (const uint8_t *)b->buffer.bytes + b->offset : NULL;
}
const void *envImage(void *user, uint32_t set, uint32_t bind)
{
  const vb200_binding *b = findBinding((const vb200_draw_state *)user, set, bind);
  return b ? &b->image : NULL;
}
const uint8_t *envPush(void *user, uint32_t offset)
{
  return ((const vb200_draw_state *)user)->pushconsts + offset;
}
void envSampleTex(void *, float u, float v, const void *img, uint64_t offs, float out[4])
{
  sampleTex(u, v, (const vb200_image *)img, offs, out);
}
void envSampleCube(void *, float x, float y, float z, const void *img, float out[4])
{
  sampleCube(x, y, z, (const vb200_image *)img, out);
}

vor::ShaderEnv makeEnv(const vb200_draw_state *s)
{
  vor::ShaderEnv e;
  e.user = (void *)s;
  e.vertex_attr = envVertexAttr;
  e.buffer_ptr = envBufferPtr;
  e.image = envImage;
  e.push_ptr = envPush;
  e.sample_tex = envSampleTex;
  e.sample_cube = envSampleCube;
  return e;
}

uint32_t getIndex(const vb200_draw_state *s, uint32_t vertexIndex, bool indexed)
{
  // rasterizer.cpp:100-119
  if(!indexed)
    return vertexIndex;
  const byte *ib = (const byte *)s->ib.buffer.bytes + s->ib.offset;
  if(s->ib.index_type == INDEX_UINT16)
  {
    uint16_t v;
    memcpy(&v, ib + 2 * (uint64_t)vertexIndex, 2);
    return v;
  }
  uint32_t v;
  memcpy(&v, ib + 4 * (uint64_t)vertexIndex, 4);
  return v;
}

inline float clamp01(float in)
{
  return in > 1.0f ? 1.0f : (in < 0.0f ? 0.0f : in);
}

struct Counters
{
  uint64_t tris_in = 0, tris_out = 0, draws = 0, pixels_tested = 0, pixels_written = 0, depth_passed = 0;
} g_counters;

}    // namespace

extern "C" {

// option "extended_spirv": mirror of vb200_set_option for the extended opcode set
__attribute__((visibility("default"))) int vor_set_option(const char *name, int64_t value)
{
  if(name && !strcmp(name, "extended_spirv"))
  {
    vor::set_extended(value != 0);
    return 0;
  }
  return 1;
}

__attribute__((visibility("default"))) int vor_init(int)
{
  return 0;
}

__attribute__((visibility("default"))) const char *vor_last_error(void)
{
  return g_err.c_str();
}

__attribute__((visibility("default"))) void *vor_shader_create(const uint32_t *code, size_t words)
{
  return vor::compile(code, words, &g_err);
}
__attribute__((visibility("default"))) void *vor_shader_entry(void *shader, const char *name)
{
  return shader ? (void *)vor::find_entry((vor::Module *)shader, name) : NULL;
}
__attribute__((visibility("default"))) void vor_shader_destroy(void *shader)
{
  if(shader)
    vor::destroy((vor::Module *)shader);
}
__attribute__((visibility("default"))) int vor_entry_stage(void *entry)
{
  return vor::entry_stage((const vor::Entry *)entry);
}

// direct access to the wrappers, for shader-stage unit tests
__attribute__((visibility("default"))) int vor_run_vertex(const vb200_draw_state *s, void *entry,
                                                          uint32_t vertexIndex, float out[44])
{
  vor::ShaderEnv env = makeEnv(s);
  vor::run_vertex((const vor::Entry *)entry, env, vertexIndex, out);
  return 0;
}
__attribute__((visibility("default"))) int vor_run_fragment(const vb200_draw_state *s, void *entry,
                                                            const float bary[4], const float tri[132],
                                                            float out[4])
{
  vor::ShaderEnv env = makeEnv(s);
  vor::run_fragment((const vor::Entry *)entry, env, 0.0f, bary, tri, out);
  return 0;
}

__attribute__((visibility("default"))) int vor_clear_color(const vb200_image *target, const float rgba[4])
{
  // rasterizer.cpp:332-361
  byte *bits = (byte *)target->pixels;
  const uint32_t w = target->width, h = target->height, bpp = target->bytes_per_pixel;
  byte eval[4];
  eval[2] = byte(rgba[0] * 255.0f);
  eval[1] = byte(rgba[1] * 255.0f);
  eval[0] = byte(rgba[2] * 255.0f);
  eval[3] = byte(rgba[3] * 255.0f);
  if(bpp == 1)
    memset(bits, eval[2], (size_t)w * h);
  else if(bpp == 4)
    for(size_t i = 0; i < (size_t)w * h; i++)
      memcpy(bits + i * 4, eval, 4);
  return 0;
}

__attribute__((visibility("default"))) int vor_clear_depth(const vb200_image *target, float depth)
{
  // rasterizer.cpp:312-330
  if(target->bytes_per_pixel != 4)
  {
    g_err = "depth clear needs bpp 4";
    return VB200_ERR_INVALID;
  }
  byte *bits = (byte *)target->pixels;
  for(size_t i = 0; i < (size_t)target->width * target->height; i++)
    memcpy(bits + i * 4, &depth, 4);
  return 0;
}

__attribute__((visibility("default"))) int vor_sample(const vb200_image *tex, int cube, uint64_t byteOffs,
                                                      const float *uvw, float *out, size_t count)
{
  for(size_t i = 0; i < count; i++)
  {
    if(cube)
      sampleCube(uvw[3 * i], uvw[3 * i + 1], uvw[3 * i + 2], tex, out + 4 * i);
    else
      sampleTex(uvw[2 * i], uvw[2 * i + 1], tex, byteOffs, out + 4 * i);
  }
  return 0;
}

__attribute__((visibility("default"))) int vor_draw(const vb200_draw_state *s, int numVerts, uint32_t first,
                                                    int indexedI)
{
  const bool indexed = indexedI != 0;
  const vb200_pipeline *pipe = s->pipeline;
  const vor::Entry *vs = (const vor::Entry *)pipe->vs;
  const vor::Entry *fs = (const vor::Entry *)pipe->fs;
  if(!vs || !fs || !s->color.pixels)
  {
    g_err = "draw without vs/fs/colour target";
    return VB200_ERR_INVALID;
  }
  vor::ShaderEnv env = makeEnv(s);

  const uint32_t w = s->color.width, h = s->color.height;

  // ---- vertex id sequence of every triangle corner (rasterizer.cpp:128-232)
  std::vector<uint32_t> corners;
  if(pipe->topology == TOPOLOGY_TRIANGLE_LIST)
  {
    int lastVert = numVerts - 3;
    uint32_t vi = first;
    for(int v = 0; v <= lastVert; v += 3)
    {
      corners.push_back(vi++);
      corners.push_back(vi++);
      corners.push_back(vi++);
    }
  }
  else if(pipe->topology == TOPOLOGY_TRIANGLE_STRIP)
  {
    if(numVerts < 3)
    {
      g_err = "strip needs >= 3 vertices (assert rasterizer.cpp:152)";
      return VB200_ERR_INVALID;
    }
    // N,N+1,N+2 / N+2,N+1,N+3 / N+2,N+3,N+4 / ...
    int ntri = numVerts - 2;
    for(int t = 0; t < ntri; t++)
    {
      uint32_t base = first + (uint32_t)t;
      if((t & 1) == 0)
      {
        corners.push_back(base);
        corners.push_back(base + 1);
        corners.push_back(base + 2);
      }
      else
      {
        corners.push_back(base + 1);
        corners.push_back(base);
        corners.push_back(base + 2);
      }
    }
  }
  else
  {
    printf("Unsupported primitive topology!\n");    // rasterizer.cpp:235 — draws nothing
  }

  // ---- ShadeVerts: VS per corner. The reference re-shades every index (no reuse); the VS is a
  // pure function of the index so a small memo is result-neutral.
  const size_t ncorner = corners.size();
  std::vector<Vertex> shaded(ncorner);
  {
    const size_t MEMO = 4096;
    std::vector<uint32_t> memoIdx(MEMO, 0xffffffffu);
    std::vector<uint8_t> memoValid(MEMO, 0);
    std::vector<Vertex> memo(MEMO);
    for(size_t c = 0; c < ncorner; c++)
    {
      uint32_t idx = getIndex(s, corners[c], indexed);
      size_t slot = idx % MEMO;
      if(!memoValid[slot] || memoIdx[slot] != idx)
      {
        memset(&memo[slot], 0, sizeof(Vertex));
        vor::run_vertex(vs, env, idx, (float *)&memo[slot]);
        memoIdx[slot] = idx;
        memoValid[slot] = 1;
      }
      shaded[c] = memo[slot];
    }
  }

  // ---- ToWindow (rasterizer.cpp:248-249)
  std::vector<int> winx(ncorner), winy(ncorner);
  for(size_t c = 0; c < ncorner; c++)
  {
    const float *p = shaded[c].position;
    winx[c] = int((p[0] / p[3] + 1.0f) * 0.5f * w);
    winy[c] = int((p[1] * -1.0f / p[3] + 1.0f) * 0.5f * h);
  }

  byte *bits = (byte *)s->color.pixels;
  const uint32_t bpp = s->color.bytes_per_pixel;
  float *depthbits = (float *)s->depth.pixels;

  for(size_t i = 0; i + 2 < ncorner; i += 3)
  {
    g_counters.tris_in++;
    const int ax = winx[i], ay = winy[i], bx = winx[i + 1], by = winy[i + 1], cx = winx[i + 2],
              cy = winy[i + 2];
    const Vertex *vsout = &shaded[i];

    // rasterizer.cpp:272-275, :395-424
    int area2 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
    if(area2 == 0)
      continue;
    int area2_flipped = area2;
    int barymul = 1;
    if(pipe->front_face == FRONT_FACE_CLOCKWISE)
    {
      barymul *= -1;
      area2_flipped *= -1;
    }
    if(area2_flipped > 0 && (pipe->cull_mode & CULL_FRONT))
      continue;
    if(area2_flipped < 0)
    {
      if(pipe->cull_mode & CULL_BACK)
        continue;
      barymul *= -1;
      area2_flipped *= -1;
    }
    g_counters.tris_out++;

    // MinMax + clamp (rasterizer.cpp:255-270, :428-435)
    int minx = ax < bx ? ax : bx, maxx = ax > bx ? ax : bx;
    int miny = ay < by ? ay : by, maxy = ay > by ? ay : by;
    minx = minx < cx ? minx : cx;
    maxx = maxx > cx ? maxx : cx;
    miny = miny < cy ? miny : cy;
    maxy = maxy > cy ? maxy : cy;
    if(minx < 0)
      minx = 0;
    if(miny < 0)
      miny = 0;
    if(maxx > int(w - 1))
      maxx = int(w - 1);
    if(maxy > int(h - 1))
      maxy = int(h - 1);

    const int ABx = bx - ax, ABy = by - ay, ACx = cx - ax, ACy = cy - ay;
    const float invarea = 1.0f / float(area2_flipped);
    const float invw[3] = {1.0f / vsout[0].position[3], 1.0f / vsout[1].position[3],
                           1.0f / vsout[2].position[3]};
    const float depth[3] = {vsout[0].position[2] * invw[0], vsout[1].position[2] * invw[1],
                            vsout[2].position[2] * invw[2]};

    // ProcessTriangles over the (exclusive) bbox (rasterizer.cpp:538-691)
    for(int y = miny; y < maxy; y++)
    {
      for(int x = minx; x < maxx; x++)
      {
        g_counters.pixels_tested++;
        const int PAx = ax - x, PAy = ay - y;
        const int ux = (ACx * PAy) - (ACy * PAx);
        const int uy = (PAx * ABy) - (PAy * ABx);
        int b0 = (area2 - (ux + uy)) * barymul, b1 = ux * barymul, b2 = uy * barymul;
        if(!(b0 >= 0 && b1 >= 0 && b2 >= 0))
          continue;
        g_counters.pixels_written++;

        float n[4] = {float(b0), float(b1), float(b2), 0.0f};
        n[0] *= invarea;
        n[1] *= invarea;
        n[2] *= invarea;
        float pixdepth = n[0] * depth[0] + n[1] * depth[1] + n[2] * depth[2];

        bool passed = true;
        const size_t pidx = (size_t)y * w + x;
        if(pipe->depth_compare_op != CMP_ALWAYS && depthbits)
        {
          float curdepth = depthbits[pidx];
          switch(pipe->depth_compare_op)
          {
            case CMP_NEVER: passed = false; break;
            case CMP_LESS: passed = pixdepth < curdepth; break;
            case CMP_EQUAL: passed = pixdepth == curdepth; break;
            case CMP_LEQUAL: passed = pixdepth <= curdepth; break;
            case CMP_GREATER: passed = pixdepth > curdepth; break;
            case CMP_NOTEQUAL: passed = pixdepth != curdepth; break;
            case CMP_GEQUAL: passed = pixdepth >= curdepth; break;
          }
        }
        if(!passed)
          continue;

        n[0] *= invw[0];
        n[1] *= invw[1];
        n[2] *= invw[2];
        float invlen = 1.0f / (n[0] + n[1] + n[2]);
        n[0] *= invlen;
        n[1] *= invlen;
        n[2] *= invlen;

        float pix[4] = {0, 0, 0, 0};
        // (extended mode only, no reference counterpart: a discarded fragment writes neither colour nor depth)
        if(vor::run_fragment(fs, env, pixdepth, n, (const float *)vsout, pix))
          continue;

        byte *px = bits + pidx * bpp;
        if(pipe->blend_enable)
        {
          // rasterizer.cpp:593-672
          float existing[4] = {float(px[2]), float(px[1]), float(px[0]), 1.0f};
          existing[0] /= 255.0f;
          existing[1] /= 255.0f;
          existing[2] /= 255.0f;
          float srcFactor = 1.0f, dstFactor = 1.0f;
          switch(pipe->src_color_blend_factor)
          {
            case BF_ZERO: srcFactor = 0.0f; break;
            case BF_ONE: srcFactor = 1.0f; break;
            case BF_SRC_ALPHA: srcFactor = pix[3]; break;
            case BF_ONE_MINUS_SRC_ALPHA: srcFactor = 1.0f - pix[3]; break;
            default: break;    // "Unsupported blend factor": stays 1.0
          }
          switch(pipe->dst_color_blend_factor)
          {
            case BF_ZERO: dstFactor = 0.0f; break;
            case BF_ONE: dstFactor = 1.0f; break;
            case BF_SRC_ALPHA: dstFactor = pix[3]; break;
            case BF_ONE_MINUS_SRC_ALPHA: dstFactor = 1.0f - pix[3]; break;
            default: break;
          }
          if(pipe->color_blend_op == BLEND_OP_ADD)
          {
            float blended[4];
            for(int c = 0; c < 4; c++)
              blended[c] = srcFactor * pix[c] + dstFactor * existing[c];
            memcpy(pix, blended, sizeof(pix));
          }
          // other ops leave `blended` uninitialised in the reference (rasterizer.cpp:655-669):
          // outside the tested domain.
        }
        px[2] = byte(clamp01(pix[0]) * 255.0f);
        px[1] = byte(clamp01(pix[1]) * 255.0f);
        px[0] = byte(clamp01(pix[2]) * 255.0f);
        g_counters.depth_passed++;
        if(pipe->depth_write_enable && depthbits)
          depthbits[pidx] = pixdepth;
      }
    }
  }
  g_counters.draws++;
  return 0;
}

__attribute__((visibility("default"))) int vor_flush(void)
{
  return 0;
}

__attribute__((visibility("default"))) int vor_get_stats(vb200_stats *out)
{
  memset(out, 0, sizeof(*out));
  out->draws = g_counters.draws;
  out->triangles_in = g_counters.tris_in;
  out->triangles_out = g_counters.tris_out;
  out->fragments_covered = g_counters.pixels_written;
  out->fragments_shaded = g_counters.depth_passed;
  return 0;
}
__attribute__((visibility("default"))) void vor_reset_stats(void)
{
  g_counters = Counters();
}
}
