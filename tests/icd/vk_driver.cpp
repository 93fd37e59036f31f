/*
 * tests/icd/vk_driver.cpp — headless Vulkan application used by tests/test_icd.py.
 *
 * dlopen()s an ICD, resolves entry points through vk_icdGetInstanceProcAddr exactly like the Vulkan
 * loader would (there is no loader or WSI in the image), and issues the complete call sequence of a
 * small renderer: instance/device, buffers+images bound to host-visible memory, texture upload through
 * vkCmdCopyBufferToImage, descriptor sets, render pass, graphics pipeline, command buffer,
 * vkQueueSubmit; then reads the attachments through the mapped pointers.  The same scene is run
 * through the reference ICD (oracle/_ref/libvisor_ref.so) and the CUDA ICD
 * (integration/libvisor_b200_icd.so).  Compiled against the reference's vendored vulkan.h (v42).
 * TEST INFRASTRUCTURE.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include <chrono>
#include <map>
#include <string>
#include <vector>
#include "3rdparty/vulkan.h"

extern "C" {
struct vkd_attr { uint32_t location, format, stride, offset, binding; };
struct vkd_ubo { uint32_t set, binding; const void *data; uint64_t size, offset; };
struct vkd_tex { uint32_t set, binding; const void *data; uint32_t width, height, format, bpp, layers; };
struct vkd_draw
{
  const uint32_t *vs_code; uint32_t vs_words; const uint32_t *fs_code; uint32_t fs_words;
  uint32_t num_attrs; vkd_attr attrs[16];
  uint32_t topology, front_face, cull_mode, depth_op, depth_write, blend_enable, src_factor, dst_factor, blend_op;
  const void *vb[4]; uint64_t vb_size[4]; uint64_t vb_offset[4];
  const void *ib; uint64_t ib_size, ib_offset; uint32_t index_type;
  uint32_t num_ubos; vkd_ubo ubos[4];
  uint32_t num_tex; vkd_tex tex[4];
  uint8_t push[128]; uint32_t push_size;
  uint32_t count, first, indexed;
};
struct vkd_scene
{
  uint32_t width, height, has_depth;
  uint32_t clear_color_enable; float clear_color[4];
  uint32_t clear_depth_enable; float clear_depth;
  uint32_t num_draws; const vkd_draw *draws;
  void *color_out; void *depth_out;
};
}

namespace
{
typedef PFN_vkVoidFunction(VKAPI_PTR *PFN_gipa)(VkInstance, const char *);

struct Ctx
{
  void *lib = NULL;
  PFN_gipa gipa = NULL;
  VkDevice dev = VK_NULL_HANDLE;
  std::vector<VkDeviceMemory> mems;
  template <typename T>
  T fn(const char *name)
  {
    T f = (T)gipa(NULL, name);
    if(!f)
      fprintf(stderr, "vk_driver: missing %s\n", name);
    return f;
  }
};

#define VK(name) PFN_##name name = c.fn<PFN_##name>(#name)

struct Mapped
{
  VkDeviceMemory mem;
  void *ptr;
};

// One instance + device per ICD per process. The reference cannot survive a second vkCreateInstance
// once it has sampled a texture: InitTextureCache (texture_sampling.cpp:19-32) relinks the LRU list
// without clearing the old head/tail links, which leaves a cycle and CacheCoord spins forever.
struct Loaded
{
  void *lib;
  PFN_gipa gipa;
  VkInstance inst;
  VkDevice dev;
  VkQueue queue;
  bool serial;
};
std::map<std::string, Loaded> g_loaded;
std::vector<double> g_frame_seconds;    // vkQueueSubmit + vkQueueWaitIdle of every frame of the last vkd_run
}    // namespace

// per-frame submit times of the last vkd_run (host wall clock around vkQueueSubmit + vkQueueWaitIdle)
extern "C" __attribute__((visibility("default"))) int vkd_frame_seconds(double *out, int n)
{
  int i = 0;
  for(; i < n && i < (int)g_frame_seconds.size(); i++)
    out[i] = g_frame_seconds[i];
  return i;
}

extern "C" __attribute__((visibility("default"))) int vkd_run(const char *icd_path, const vkd_scene *sc, int frames,
                                                              int serial_reference, double *submit_seconds)
{
  Ctx c;
  const bool first = g_loaded.find(icd_path) == g_loaded.end();
  if(first)
  {
    Loaded l;
    memset(&l, 0, sizeof(l));
    l.lib = dlopen(icd_path, RTLD_NOW | RTLD_LOCAL);
    if(!l.lib)
    {
      fprintf(stderr, "vk_driver: dlopen(%s): %s\n", icd_path, dlerror());
      return -1;
    }
    l.gipa = (PFN_gipa)dlsym(l.lib, "vk_icdGetInstanceProcAddr");
    if(!l.gipa)
      return -2;
    g_loaded[icd_path] = l;
  }
  Loaded &L = g_loaded[icd_path];
  c.lib = L.lib;
  c.gipa = L.gipa;
  VK(vkCreateInstance); VK(vkEnumeratePhysicalDevices); VK(vkCreateDevice); VK(vkGetDeviceQueue);
  VK(vkCreateBuffer); VK(vkGetBufferMemoryRequirements); VK(vkAllocateMemory); VK(vkBindBufferMemory);
  VK(vkMapMemory); VK(vkCreateImage); VK(vkGetImageMemoryRequirements); VK(vkBindImageMemory);
  VK(vkCreateImageView); VK(vkCreateRenderPass); VK(vkCreateFramebuffer); VK(vkCreateShaderModule);
  VK(vkCreateDescriptorSetLayout); VK(vkCreatePipelineLayout); VK(vkCreateDescriptorPool);
  VK(vkAllocateDescriptorSets); VK(vkUpdateDescriptorSets); VK(vkCreateGraphicsPipelines);
  VK(vkCreateCommandPool); VK(vkAllocateCommandBuffers); VK(vkBeginCommandBuffer); VK(vkCmdBeginRenderPass);
  VK(vkCmdBindPipeline); VK(vkCmdBindDescriptorSets); VK(vkCmdBindVertexBuffers); VK(vkCmdBindIndexBuffer);
  VK(vkCmdPushConstants); VK(vkCmdDraw); VK(vkCmdDrawIndexed); VK(vkCmdEndRenderPass); VK(vkEndCommandBuffer);
  VK(vkQueueSubmit); VK(vkCmdCopyBufferToImage); VK(vkCmdSetViewport); VK(vkQueueWaitIdle);
  VK(vkFreeMemory); VK(vkCreateSampler);

  if(first)
  {
    VkInstanceCreateInfo ici = {VK_STRUCTURE_TYPE_INSTANCE_CREATE_INFO};
    vkCreateInstance(&ici, NULL, &L.inst);
    uint32_t npd = 1;
    VkPhysicalDevice pd;
    vkEnumeratePhysicalDevices(L.inst, &npd, &pd);
    float prio = 1.0f;
    VkDeviceQueueCreateInfo qci = {VK_STRUCTURE_TYPE_DEVICE_QUEUE_CREATE_INFO, NULL, 0, 0, 1, &prio};
    VkDeviceCreateInfo dci = {VK_STRUCTURE_TYPE_DEVICE_CREATE_INFO};
    dci.queueCreateInfoCount = 1;
    dci.pQueueCreateInfos = &qci;
    vkCreateDevice(pd, &dci, NULL, &L.dev);
    vkGetDeviceQueue(L.dev, 0, 0, &L.queue);
  }
  if(serial_reference && !L.serial)
  {
    // the reference's vkCreateInstance spawns its 7 racy workers (icd_stubs.cpp:12); joining them
    // puts it in the deterministic serial-drain mode. For the CUDA ICD it is a flush.
    // (One way only: a process that wants the threaded mode timed must run it before any serial call.)
    void (*shutdownThreads)() = (void (*)())dlsym(c.lib, "_Z21ShutdownRasterThreadsv");
    if(shutdownThreads)
      shutdownThreads();
    L.serial = true;
  }
  VkDevice dev = L.dev;
  VkQueue queue = L.queue;

  std::vector<VkDeviceMemory> allMem;
  // memory type 1 = HOST_VISIBLE|COHERENT|CACHED (mapped), type 0 = DEVICE_LOCAL (never mapped), query.cpp:246-267
  auto allocFor = [&](VkMemoryRequirements req, uint32_t memoryType = 1) -> Mapped {
    VkMemoryAllocateInfo ai = {VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO, NULL, req.size ? req.size : 16, memoryType};
    Mapped m;
    m.ptr = NULL;
    vkAllocateMemory(dev, &ai, NULL, &m.mem);
    if(memoryType == 1)
      vkMapMemory(dev, m.mem, 0, VK_WHOLE_SIZE, 0, &m.ptr);
    allMem.push_back(m.mem);
    return m;
  };
  auto makeBuffer = [&](const void *data, uint64_t size, VkBuffer *out) -> Mapped {
    VkBufferCreateInfo bi = {VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO};
    bi.size = size;
    bi.usage = VK_BUFFER_USAGE_VERTEX_BUFFER_BIT | VK_BUFFER_USAGE_INDEX_BUFFER_BIT |
               VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT | VK_BUFFER_USAGE_TRANSFER_SRC_BIT;
    vkCreateBuffer(dev, &bi, NULL, out);
    VkMemoryRequirements req;
    vkGetBufferMemoryRequirements(dev, *out, &req);
    Mapped m = allocFor(req);
    vkBindBufferMemory(dev, *out, m.mem, 0);
    if(data)
      memcpy(m.ptr, data, size);
    return m;
  };
  auto makeImage = [&](uint32_t w, uint32_t h, VkFormat fmt, uint32_t layers, VkImageUsageFlags usage, VkImage *out,
                       VkImageView *view) -> Mapped {
    VkImageCreateInfo ii = {VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO};
    ii.imageType = VK_IMAGE_TYPE_2D;
    ii.format = fmt;
    ii.extent = {w, h, 1};
    ii.mipLevels = 1;
    ii.arrayLayers = layers;
    ii.samples = VK_SAMPLE_COUNT_1_BIT;
    ii.tiling = VK_IMAGE_TILING_LINEAR;
    ii.usage = usage;
    if(layers == 6)
      ii.flags = VK_IMAGE_CREATE_CUBE_COMPATIBLE_BIT;
    vkCreateImage(dev, &ii, NULL, out);
    VkMemoryRequirements req;
    vkGetImageMemoryRequirements(dev, *out, &req);
    // sampled textures live in DEVICE_LOCAL memory and are filled by vkCmdCopyBufferToImage, as in a real
    // application; attachments are host-visible because the test reads them back through the mapping
    Mapped m = allocFor(req, (usage & VK_IMAGE_USAGE_SAMPLED_BIT) ? 0 : 1);
    vkBindImageMemory(dev, *out, m.mem, 0);
    VkImageViewCreateInfo vi = {VK_STRUCTURE_TYPE_IMAGE_VIEW_CREATE_INFO};
    vi.image = *out;
    vi.viewType = layers == 6 ? VK_IMAGE_VIEW_TYPE_CUBE : VK_IMAGE_VIEW_TYPE_2D;
    vi.format = fmt;
    vi.subresourceRange = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 1, 0, layers};
    vkCreateImageView(dev, &vi, NULL, view);
    return m;
  };

  // ---- attachments
  VkImage colImg, depImg = VK_NULL_HANDLE;
  VkImageView colView, depView = VK_NULL_HANDLE;
  Mapped colMem = makeImage(sc->width, sc->height, VK_FORMAT_B8G8R8A8_UNORM, 1, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT,
                            &colImg, &colView);
  Mapped depMem = {VK_NULL_HANDLE, NULL};
  const size_t px = (size_t)sc->width * sc->height;
  memcpy(colMem.ptr, sc->color_out, px * 4);    // initial contents (matter when loadOp = LOAD)
  if(sc->has_depth)
  {
    depMem = makeImage(sc->width, sc->height, VK_FORMAT_D32_SFLOAT, 1, VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT,
                       &depImg, &depView);
    memcpy(depMem.ptr, sc->depth_out, px * 4);
  }

  VkAttachmentDescription att[2];
  memset(att, 0, sizeof(att));
  att[0].format = VK_FORMAT_B8G8R8A8_UNORM;
  att[0].samples = VK_SAMPLE_COUNT_1_BIT;
  att[0].loadOp = sc->clear_color_enable ? VK_ATTACHMENT_LOAD_OP_CLEAR : VK_ATTACHMENT_LOAD_OP_LOAD;
  att[0].storeOp = VK_ATTACHMENT_STORE_OP_STORE;
  att[1].format = VK_FORMAT_D32_SFLOAT;
  att[1].samples = VK_SAMPLE_COUNT_1_BIT;
  att[1].loadOp = sc->clear_depth_enable ? VK_ATTACHMENT_LOAD_OP_CLEAR : VK_ATTACHMENT_LOAD_OP_LOAD;
  att[1].storeOp = VK_ATTACHMENT_STORE_OP_STORE;
  VkAttachmentReference colRef = {0, VK_IMAGE_LAYOUT_COLOR_ATTACHMENT_OPTIMAL};
  VkAttachmentReference depRef = {1, VK_IMAGE_LAYOUT_DEPTH_STENCIL_ATTACHMENT_OPTIMAL};
  VkSubpassDescription sub;
  memset(&sub, 0, sizeof(sub));
  sub.pipelineBindPoint = VK_PIPELINE_BIND_POINT_GRAPHICS;
  sub.colorAttachmentCount = 1;
  sub.pColorAttachments = &colRef;
  sub.pDepthStencilAttachment = sc->has_depth ? &depRef : NULL;
  VkRenderPassCreateInfo rpi = {VK_STRUCTURE_TYPE_RENDER_PASS_CREATE_INFO};
  rpi.attachmentCount = sc->has_depth ? 2 : 1;
  rpi.pAttachments = att;
  rpi.subpassCount = 1;
  rpi.pSubpasses = &sub;
  VkRenderPass rp;
  vkCreateRenderPass(dev, &rpi, NULL, &rp);
  VkImageView fbViews[2] = {colView, depView};
  VkFramebufferCreateInfo fbi = {VK_STRUCTURE_TYPE_FRAMEBUFFER_CREATE_INFO};
  fbi.renderPass = rp;
  fbi.attachmentCount = rpi.attachmentCount;
  fbi.pAttachments = fbViews;
  fbi.width = sc->width;
  fbi.height = sc->height;
  fbi.layers = 1;
  VkFramebuffer fb;
  vkCreateFramebuffer(dev, &fbi, NULL, &fb);

  VkCommandPoolCreateInfo cpi = {VK_STRUCTURE_TYPE_COMMAND_POOL_CREATE_INFO};
  VkCommandPool pool;
  vkCreateCommandPool(dev, &cpi, NULL, &pool);
  VkCommandBufferAllocateInfo cai = {VK_STRUCTURE_TYPE_COMMAND_BUFFER_ALLOCATE_INFO, NULL, pool,
                                     VK_COMMAND_BUFFER_LEVEL_PRIMARY, 1};
  VkCommandBuffer uploadCb, cb;
  vkAllocateCommandBuffers(dev, &cai, &uploadCb);
  vkAllocateCommandBuffers(dev, &cai, &cb);
  VkCommandBufferBeginInfo cbi = {VK_STRUCTURE_TYPE_COMMAND_BUFFER_BEGIN_INFO};

  // ---- per-draw objects
  struct DrawObjs
  {
    VkPipeline pipe;
    VkPipelineLayout layout;
    VkBuffer vb[4], ib;
    std::vector<std::pair<uint32_t, VkDescriptorSet>> sets;
  };
  std::vector<DrawObjs> objs(sc->num_draws);
  VkDescriptorPoolSize psz = {VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, 64};
  VkDescriptorPoolCreateInfo dpi = {VK_STRUCTURE_TYPE_DESCRIPTOR_POOL_CREATE_INFO, NULL, 0, 64, 1, &psz};
  VkDescriptorPool dpool;
  vkCreateDescriptorPool(dev, &dpi, NULL, &dpool);
  VkSamplerCreateInfo sci = {VK_STRUCTURE_TYPE_SAMPLER_CREATE_INFO};
  VkSampler sampler;
  vkCreateSampler(dev, &sci, NULL, &sampler);

  vkBeginCommandBuffer(uploadCb, &cbi);
  for(uint32_t di = 0; di < sc->num_draws; di++)
  {
    const vkd_draw &d = sc->draws[di];
    DrawObjs &o = objs[di];
    // a draw that differs from an earlier one only in its vertex/index range uses that draw's objects, as an
    // application drawing a mesh chunk by chunk would (same buffers, descriptor sets and pipeline)
    {
      uint32_t same = di;
      for(uint32_t dj = 0; dj < di && same == di; dj++)
        if(memcmp(&sc->draws[dj], &d, offsetof(vkd_draw, count)) == 0)
          same = dj;
      if(same != di)
      {
        o = objs[same];
        continue;
      }
    }
    memset(o.vb, 0, sizeof(o.vb));
    o.ib = VK_NULL_HANDLE;
    for(int i = 0; i < 4; i++)
      if(d.vb[i])
        makeBuffer(d.vb[i], d.vb_size[i], &o.vb[i]);
    if(d.ib)
      makeBuffer(d.ib, d.ib_size, &o.ib);

    // descriptor sets: one layout + one vkAllocateDescriptorSets call per set (descriptors.cpp:58-63
    // writes index 0 only)
    uint32_t maxBind[8];
    bool used[8];
    memset(maxBind, 0, sizeof(maxBind));
    memset(used, 0, sizeof(used));
    for(uint32_t u = 0; u < d.num_ubos; u++)
    {
      used[d.ubos[u].set] = true;
      maxBind[d.ubos[u].set] = std::max(maxBind[d.ubos[u].set], d.ubos[u].binding);
    }
    for(uint32_t t = 0; t < d.num_tex; t++)
    {
      used[d.tex[t].set] = true;
      maxBind[d.tex[t].set] = std::max(maxBind[d.tex[t].set], d.tex[t].binding);
    }
    std::vector<VkDescriptorSetLayout> layouts;
    VkDescriptorSet setOf[8];
    for(uint32_t s = 0; s < 8; s++)
    {
      if(!used[s])
        continue;
      std::vector<VkDescriptorSetLayoutBinding> lb(maxBind[s] + 1);
      for(uint32_t b = 0; b <= maxBind[s]; b++)
        lb[b] = {b, VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, 1, VK_SHADER_STAGE_ALL_GRAPHICS, NULL};
      VkDescriptorSetLayoutCreateInfo li = {VK_STRUCTURE_TYPE_DESCRIPTOR_SET_LAYOUT_CREATE_INFO, NULL, 0,
                                            (uint32_t)lb.size(), lb.data()};
      VkDescriptorSetLayout dsl;
      vkCreateDescriptorSetLayout(dev, &li, NULL, &dsl);
      layouts.push_back(dsl);
      VkDescriptorSetAllocateInfo ai = {VK_STRUCTURE_TYPE_DESCRIPTOR_SET_ALLOCATE_INFO, NULL, dpool, 1, &dsl};
      vkAllocateDescriptorSets(dev, &ai, &setOf[s]);
      o.sets.push_back({s, setOf[s]});
    }
    for(uint32_t u = 0; u < d.num_ubos; u++)
    {
      VkBuffer ub;
      makeBuffer(d.ubos[u].data, d.ubos[u].size, &ub);
      VkDescriptorBufferInfo bi = {ub, d.ubos[u].offset, d.ubos[u].size - d.ubos[u].offset};
      VkWriteDescriptorSet w = {VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, NULL, setOf[d.ubos[u].set], d.ubos[u].binding,
                                0, 1, VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER, NULL, &bi, NULL};
      vkUpdateDescriptorSets(dev, 1, &w, 0, NULL);
    }
    for(uint32_t t = 0; t < d.num_tex; t++)
    {
      const vkd_tex &tx = d.tex[t];
      VkImage ti;
      VkImageView tv;
      makeImage(tx.width, tx.height, (VkFormat)tx.format, tx.layers,
                VK_IMAGE_USAGE_SAMPLED_BIT | VK_IMAGE_USAGE_TRANSFER_DST_BIT, &ti, &tv);
      // staging buffer + one vkCmdCopyBufferToImage per layer (cmd_exec.cpp:143-173)
      const uint64_t layerBytes = (uint64_t)tx.width * tx.height * tx.bpp;
      VkBuffer staging;
      makeBuffer(tx.data, layerBytes * tx.layers, &staging);
      for(uint32_t l = 0; l < tx.layers; l++)
      {
        VkBufferImageCopy r;
        memset(&r, 0, sizeof(r));
        r.bufferOffset = layerBytes * l;
        r.imageSubresource = {VK_IMAGE_ASPECT_COLOR_BIT, 0, l, 1};
        r.imageExtent = {tx.width, tx.height, 1};
        vkCmdCopyBufferToImage(uploadCb, staging, ti, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &r);
      }
      VkDescriptorImageInfo ii = {sampler, tv, VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL};
      VkWriteDescriptorSet w = {VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET, NULL, setOf[tx.set], tx.binding, 0, 1,
                                VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER, &ii, NULL, NULL};
      vkUpdateDescriptorSets(dev, 1, &w, 0, NULL);
    }
    VkPushConstantRange pcr = {VK_SHADER_STAGE_ALL_GRAPHICS, 0, 128};
    VkPipelineLayoutCreateInfo pli = {VK_STRUCTURE_TYPE_PIPELINE_LAYOUT_CREATE_INFO, NULL, 0, (uint32_t)layouts.size(),
                                      layouts.data(), 1, &pcr};
    vkCreatePipelineLayout(dev, &pli, NULL, &o.layout);

    // ---- pipeline
    VkShaderModule vsm, fsm;
    VkShaderModuleCreateInfo smi = {VK_STRUCTURE_TYPE_SHADER_MODULE_CREATE_INFO, NULL, 0, d.vs_words * 4ull, d.vs_code};
    if(vkCreateShaderModule(dev, &smi, NULL, &vsm) != VK_SUCCESS)
      return -10;
    smi.codeSize = d.fs_words * 4ull;
    smi.pCode = d.fs_code;
    if(vkCreateShaderModule(dev, &smi, NULL, &fsm) != VK_SUCCESS)
      return -11;
    VkPipelineShaderStageCreateInfo stages[2];
    memset(stages, 0, sizeof(stages));
    stages[0].sType = stages[1].sType = VK_STRUCTURE_TYPE_PIPELINE_SHADER_STAGE_CREATE_INFO;
    stages[0].stage = VK_SHADER_STAGE_VERTEX_BIT;
    stages[0].module = vsm;
    stages[0].pName = "main";
    stages[1].stage = VK_SHADER_STAGE_FRAGMENT_BIT;
    stages[1].module = fsm;
    stages[1].pName = "main";
    std::vector<VkVertexInputBindingDescription> vbd;
    std::vector<VkVertexInputAttributeDescription> vad;
    for(uint32_t a = 0; a < d.num_attrs; a++)
    {
      bool have = false;
      for(auto &b : vbd)
        have |= b.binding == d.attrs[a].binding;
      if(!have)
        vbd.push_back({d.attrs[a].binding, d.attrs[a].stride, VK_VERTEX_INPUT_RATE_VERTEX});
      vad.push_back({d.attrs[a].location, d.attrs[a].binding, (VkFormat)d.attrs[a].format, d.attrs[a].offset});
    }
    VkPipelineVertexInputStateCreateInfo vis = {VK_STRUCTURE_TYPE_PIPELINE_VERTEX_INPUT_STATE_CREATE_INFO, NULL, 0,
                                                (uint32_t)vbd.size(), vbd.data(), (uint32_t)vad.size(), vad.data()};
    VkPipelineInputAssemblyStateCreateInfo ias = {VK_STRUCTURE_TYPE_PIPELINE_INPUT_ASSEMBLY_STATE_CREATE_INFO, NULL, 0,
                                                  (VkPrimitiveTopology)d.topology, VK_FALSE};
    VkViewport vp = {0, 0, (float)sc->width, (float)sc->height, 0, 1};
    VkRect2D scissor = {{0, 0}, {sc->width, sc->height}};
    VkPipelineViewportStateCreateInfo vps = {VK_STRUCTURE_TYPE_PIPELINE_VIEWPORT_STATE_CREATE_INFO, NULL, 0, 1, &vp, 1,
                                             &scissor};
    VkPipelineRasterizationStateCreateInfo rs = {VK_STRUCTURE_TYPE_PIPELINE_RASTERIZATION_STATE_CREATE_INFO};
    rs.polygonMode = VK_POLYGON_MODE_FILL;
    rs.cullMode = d.cull_mode;
    rs.frontFace = (VkFrontFace)d.front_face;
    rs.lineWidth = 1.0f;
    VkPipelineMultisampleStateCreateInfo ms = {VK_STRUCTURE_TYPE_PIPELINE_MULTISAMPLE_STATE_CREATE_INFO};
    ms.rasterizationSamples = VK_SAMPLE_COUNT_1_BIT;
    VkPipelineDepthStencilStateCreateInfo ds = {VK_STRUCTURE_TYPE_PIPELINE_DEPTH_STENCIL_STATE_CREATE_INFO};
    ds.depthTestEnable = VK_TRUE;
    ds.depthWriteEnable = d.depth_write ? VK_TRUE : VK_FALSE;
    ds.depthCompareOp = (VkCompareOp)d.depth_op;
    VkPipelineColorBlendAttachmentState ba;
    memset(&ba, 0, sizeof(ba));
    ba.blendEnable = d.blend_enable ? VK_TRUE : VK_FALSE;
    ba.srcColorBlendFactor = (VkBlendFactor)d.src_factor;
    ba.dstColorBlendFactor = (VkBlendFactor)d.dst_factor;
    ba.colorBlendOp = (VkBlendOp)d.blend_op;
    ba.srcAlphaBlendFactor = VK_BLEND_FACTOR_ONE;
    ba.dstAlphaBlendFactor = VK_BLEND_FACTOR_ZERO;
    ba.colorWriteMask = 0xf;
    VkPipelineColorBlendStateCreateInfo cbs = {VK_STRUCTURE_TYPE_PIPELINE_COLOR_BLEND_STATE_CREATE_INFO};
    cbs.attachmentCount = 1;
    cbs.pAttachments = &ba;
    VkGraphicsPipelineCreateInfo gpi = {VK_STRUCTURE_TYPE_GRAPHICS_PIPELINE_CREATE_INFO};
    gpi.stageCount = 2;
    gpi.pStages = stages;
    gpi.pVertexInputState = &vis;
    gpi.pInputAssemblyState = &ias;
    gpi.pViewportState = &vps;
    gpi.pRasterizationState = &rs;
    gpi.pMultisampleState = &ms;
    gpi.pDepthStencilState = &ds;    // the reference runs the test iff op != ALWAYS (rasterizer.cpp:562)
    gpi.pColorBlendState = &cbs;
    gpi.layout = o.layout;
    gpi.renderPass = rp;
    vkCreateGraphicsPipelines(dev, VK_NULL_HANDLE, 1, &gpi, NULL, &o.pipe);
  }
  vkEndCommandBuffer(uploadCb);
  VkSubmitInfo usi = {VK_STRUCTURE_TYPE_SUBMIT_INFO};
  usi.commandBufferCount = 1;
  usi.pCommandBuffers = &uploadCb;
  vkQueueSubmit(queue, 1, &usi, VK_NULL_HANDLE);

  // ---- the frame
  vkBeginCommandBuffer(cb, &cbi);
  VkClearValue clears[2];
  memset(clears, 0, sizeof(clears));
  int nclear = 0;
  if(sc->clear_color_enable)
    memcpy(clears[nclear++].color.float32, sc->clear_color, 16);
  if(sc->has_depth && sc->clear_depth_enable)
    clears[nclear++].depthStencil.depth = sc->clear_depth;
  VkRenderPassBeginInfo rbi = {VK_STRUCTURE_TYPE_RENDER_PASS_BEGIN_INFO, NULL, rp, fb,
                               {{0, 0}, {sc->width, sc->height}}, (uint32_t)nclear, clears};
  vkCmdBeginRenderPass(cb, &rbi, VK_SUBPASS_CONTENTS_INLINE);
  VkViewport vp = {0, 0, (float)sc->width, (float)sc->height, 0, 1};
  vkCmdSetViewport(cb, 0, 1, &vp);
  for(uint32_t di = 0; di < sc->num_draws; di++)
  {
    const vkd_draw &d = sc->draws[di];
    DrawObjs &o = objs[di];
    vkCmdBindPipeline(cb, VK_PIPELINE_BIND_POINT_GRAPHICS, o.pipe);
    for(auto &s : o.sets)
      vkCmdBindDescriptorSets(cb, VK_PIPELINE_BIND_POINT_GRAPHICS, o.layout, s.first, 1, &s.second, 0, NULL);
    for(uint32_t i = 0; i < 4; i++)
      if(o.vb[i])
      {
        VkDeviceSize off = d.vb_offset[i];
        vkCmdBindVertexBuffers(cb, i, 1, &o.vb[i], &off);
      }
    if(o.ib)
      vkCmdBindIndexBuffer(cb, o.ib, d.ib_offset, (VkIndexType)d.index_type);
    if(d.push_size)
      vkCmdPushConstants(cb, o.layout, VK_SHADER_STAGE_ALL_GRAPHICS, 0, d.push_size, d.push);
    if(d.indexed)
      vkCmdDrawIndexed(cb, d.count, 1, d.first, 0, 0);
    else
      vkCmdDraw(cb, d.count, 1, d.first, 0);
  }
  vkCmdEndRenderPass(cb);
  vkEndCommandBuffer(cb);

  VkSubmitInfo si = {VK_STRUCTURE_TYPE_SUBMIT_INFO};
  si.commandBufferCount = 1;
  si.pCommandBuffers = &cb;
  double best = 1e30;
  g_frame_seconds.clear();
  for(int f = 0; f < frames; f++)
  {
    if(f > 0 && !sc->clear_color_enable)
    {
      memcpy(colMem.ptr, sc->color_out, px * 4);
      if(sc->has_depth)
        memcpy(depMem.ptr, sc->depth_out, px * 4);
    }
    auto t0 = std::chrono::steady_clock::now();
    VkResult r = vkQueueSubmit(queue, 1, &si, VK_NULL_HANDLE);
    vkQueueWaitIdle(queue);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if(r != VK_SUCCESS)
      return -20;
    best = dt < best ? dt : best;
    g_frame_seconds.push_back(dt);
  }
  if(submit_seconds)
    *submit_seconds = best;
  memcpy(sc->color_out, colMem.ptr, px * 4);
  if(sc->has_depth)
    memcpy(sc->depth_out, depMem.ptr, px * 4);
  for(VkDeviceMemory m : allMem)
    vkFreeMemory(dev, m, NULL);
  // the ICD stays loaded (its global rasterizer / CUDA state is process-wide)
  return 0;
}
