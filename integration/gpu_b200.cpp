/*
 * integration/gpu_b200.cpp — the reference-side binding: visor's internal "GPU" operator API
 * (gpu.h:59-70, spirv_compile.h:3-10) implemented on top of the C-ABI of libvisor_b200.so.
 *
 * Drop-in point: link this file INSTEAD OF rasterizer.cpp, texture_sampling.cpp and spirv_compile.cpp.
 * Every other translation unit of the ICD (icd_interface.cpp, cmd_record.cpp, shaders.cpp, images.cpp,
 * descriptors.cpp, ...) is compiled unmodified; they only ever call the functions defined here
 * (cmd_exec.cpp:48,59,133,140; icd_stubs.cpp:12-13,20; shaders.cpp:11,23,81,85).
 *
 * Compiled against the reference's own headers (-I<visor> ; on Linux with -include integration/linux/shim.h).
 * It contains no rasterisation, sampling or shader code: it only re-packs the reference's host structs
 * (GPUState, VkPipeline_T, VkImage_T, VkBuffer_T, VkDescriptorSet_T) into the ABI's PODs.
 */
#include "precompiled.h"
#include "gpu.h"
#include "spirv_compile.h"

#include "../include/visor_b200.h"

// LLVMFunction is opaque to every caller (only ever a pointer, precompiled.h:49,113)
struct LLVMFunction
{
  vb200_shader *shader;
};

static void reportError(const char *what)
{
  // the reference has no error channel below vkQueueSubmit (all boundary functions return void);
  // like its own printf diagnostics (rasterizer.cpp:235,624) failures go to stdout
  printf("visor_b200: %s: %s\n", what, vb200_last_error());
}

static void toImage(const VkImage_T *in, vb200_image &out)
{
  out.pixels = in->pixels;
  out.width = in->extent.width;
  out.height = in->extent.height;
  out.depth = in->extent.depth;
  out.image_type = (uint32_t)in->imageType;
  out.format = (uint32_t)in->format;
  out.array_layers = in->arrayLayers;
  out.mip_levels = in->mipLevels;
  out.bytes_per_pixel = in->bytesPerPixel;
}

// ---- spirv_compile.h -------------------------------------------------------------------------
// Which GPU, and whether this process is one rank of a sort-first group, comes from the environment the way
// launchers (torchrun, mpirun, a shell loop) already provide it:
//   VISOR_B200_DEVICE, else LOCAL_RANK      CUDA device of this process (default 0)
//   VISOR_B200_SESSION + RANK + WORLD_SIZE  join a sort-first group of WORLD_SIZE processes running the same
//                                           application: screen tiles are split across the ranks, every rank's
//                                           framebuffer memory receives the whole image at vkQueueSubmit
static int envInt(const char *name, int fallback)
{
  const char *v = getenv(name);
  return v && *v ? atoi(v) : fallback;
}

void InitLLVM()
{
  const int device = envInt("VISOR_B200_DEVICE", envInt("LOCAL_RANK", 0));
  const char *session = getenv("VISOR_B200_SESSION");
  const int world = envInt("WORLD_SIZE", 1);
  if(session && *session && world > 1)
  {
    if(vb200_mgpu_init(envInt("RANK", 0), world, device, session) != VB200_OK ||
       vb200_set_option("mgpu_mirrors", 1) != VB200_OK)
      reportError("vb200_mgpu_init");
  }
  else if(vb200_init(device) != VB200_OK)
    reportError("vb200_init");
}

void ShutdownLLVM()
{
}

LLVMFunction *CompileFunction(const uint32_t *pCode, size_t codeSize)
{
  vb200_shader *s = vb200_shader_create(pCode, codeSize);
  if(!s)
  {
    reportError("CompileFunction");
    return NULL;    // -> VK_ERROR_DEVICE_LOST (shaders.cpp:13-14)
  }
  LLVMFunction *f = new LLVMFunction;
  f->shader = s;
  return f;
}

// VkPipeline_T::vs/fs are typed as CPU function pointers (precompiled.h:53-55,131-132) but only the
// rasterizer ever calls them; shaders.cpp:81,85 just stores what we return. We return the entry handle.
Shader GetFuncPointer(LLVMFunction *func, const char *name)
{
  return (Shader)vb200_shader_entry(func->shader, name);
}

void DestroyFunction(LLVMFunction *func)
{
  if(!func)
    return;
  vb200_shader_destroy(func->shader);
  delete func;
}

// ---- gpu.h -----------------------------------------------------------------------------------
void InitRasterThreads()
{
  // no worker threads: the "GPU" is a CUDA stream (created by vb200_init)
}

void ShutdownRasterThreads()
{
  vb200_flush();
}

void InitTextureCache()
{
  // the 4x4 texel LRU (texture_sampling.cpp:5-32) is replaced by the GPU's read-only L1 path
}

void ClearTarget(VkImage target, const VkClearColorValue &col)
{
  vb200_image im;
  toImage(target, im);
  if(vb200_clear_color(&im, col.float32) != VB200_OK)
    reportError("ClearTarget");
}

void ClearTarget(VkImage target, const VkClearDepthStencilValue &col)
{
  vb200_image im;
  toImage(target, im);
  if(vb200_clear_depth(&im, col.depth) != VB200_OK)
    reportError("ClearTarget");
}

static void addBindings(const GPUState &state, const vb200_entry *e, std::vector<vb200_binding> &out)
{
  const int n = vb200_entry_num_resources(e);
  for(int i = 0; i < n; i++)
  {
    uint32_t set, binding, isImage;
    vb200_entry_resource(e, i, &set, &binding, &isImage);
    if(set >= 8 || !state.sets[set])
      continue;
    const VkDescriptorSet_T::Bind &b = state.sets[set]->binds[binding];
    vb200_binding d;
    memset(&d, 0, sizeof(d));
    d.set = set;
    d.binding = binding;
    d.type = (uint32_t)b.type;
    d.is_image = isImage;
    if(isImage)
      toImage(b.data.imageInfo.imageView->image, d.image);    // GetDescriptorImage (spirv_compile.cpp:560-564)
    else
    {
      d.buffer.bytes = b.data.bufferInfo.buffer->bytes;    // GetDescriptorBufferPointer (:552-558)
      d.buffer.size = b.data.bufferInfo.buffer->size;
      d.offset = b.data.bufferInfo.offset;
    }
    out.push_back(d);
  }
}

void DrawTriangles(const GPUState &state, int numVerts, uint32_t first, bool indexed)
{
  const VkPipeline_T *p = state.pipeline;
  if(!p || !state.col[0])
    return;

  vb200_pipeline pl;
  memset(&pl, 0, sizeof(pl));
  for(int i = 0; i < 16; i++)
  {
    pl.vattrs[i].format = (uint32_t)p->vattrs[i].format;
    pl.vattrs[i].stride = p->vattrs[i].stride;
    pl.vattrs[i].offset = p->vattrs[i].offset;
    pl.vattrs[i].vb = p->vattrs[i].vb;
  }
  pl.topology = (uint32_t)p->topology;
  pl.front_face = (uint32_t)p->frontFace;
  pl.cull_mode = (uint32_t)p->cullMode;
  pl.depth_compare_op = (uint32_t)p->depthCompareOp;
  pl.depth_write_enable = p->depthWriteEnable ? 1 : 0;
  pl.blend_enable = p->blend.blendEnable ? 1 : 0;
  pl.src_color_blend_factor = (uint32_t)p->blend.srcColorBlendFactor;
  pl.dst_color_blend_factor = (uint32_t)p->blend.dstColorBlendFactor;
  pl.color_blend_op = (uint32_t)p->blend.colorBlendOp;
  pl.vs = (const vb200_entry *)p->vs;
  pl.fs = (const vb200_entry *)p->fs;

  vb200_draw_state s;
  memset(&s, 0, sizeof(s));
  if(indexed && state.ib.buffer)
  {
    s.ib.buffer.bytes = state.ib.buffer->bytes;
    s.ib.buffer.size = state.ib.buffer->size;
    s.ib.offset = state.ib.offset;
    s.ib.index_type = (uint32_t)state.ib.indexType;
  }
  for(int i = 0; i < 4; i++)
    if(state.vbs[i].buffer)
    {
      s.vbs[i].buffer.bytes = state.vbs[i].buffer->bytes;
      s.vbs[i].buffer.size = state.vbs[i].buffer->size;
      s.vbs[i].offset = state.vbs[i].offset;
    }
  toImage(state.col[0], s.color);
  if(state.depth)
    toImage(state.depth, s.depth);
  s.pipeline = &pl;
  std::vector<vb200_binding> binds;
  if(pl.vs)
    addBindings(state, pl.vs, binds);
  if(pl.fs)
    addBindings(state, pl.fs, binds);
  s.bindings = binds.empty() ? NULL : binds.data();
  s.num_bindings = (uint32_t)binds.size();
  memcpy(s.pushconsts, state.pushconsts, sizeof(s.pushconsts));

  if(vb200_draw(&s, numVerts, first, indexed ? 1 : 0) != VB200_OK)
    reportError("DrawTriangles");
}

// The JIT'd shaders of the reference call these by name; here sampling happens inside the CUDA
// kernels. They stay exported (gpu.h:64-67) and evaluate one sample on the device.
extern "C" __declspec(dllexport) void sample_tex_wrapped(float u, float v, VkImage tex, VkDeviceSize byteOffs,
                                                         float4 &out)
{
  vb200_image im;
  toImage(tex, im);
  float uv[2] = {u, v};
  if(vb200_sample(&im, 0, byteOffs, uv, out.v, 1) != VB200_OK)
    reportError("sample_tex_wrapped");
}

extern "C" __declspec(dllexport) void sample_cube_wrapped(float x, float y, float z, VkImage tex, float4 &out)
{
  vb200_image im;
  toImage(tex, im);
  float d[3] = {x, y, z};
  if(vb200_sample(&im, 1, 0, d, out.v, 1) != VB200_OK)
    reportError("sample_cube_wrapped");
}
