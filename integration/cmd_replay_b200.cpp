/*
 * integration/cmd_replay_b200.cpp — command replay for the CUDA path; link INSTEAD OF cmd_exec.cpp.
 *
 * Same wire format (commands.h, recorded by the unmodified cmd_record.cpp/cmd_alloc.cpp) and the same
 * state tracking as VkCommandBuffer_T::execute (cmd_exec.cpp:15-185); what changes is what the
 * north star asks for: draws and clears go through the thin C-ABI asynchronously, and vkQueueSubmit
 * ends with ONE flush that waits for the stream and makes the attachments host-visible — the
 * reference's submit is synchronous and its memory coherent (cmd_exec.cpp:187-201, memory.cpp:36-41),
 * so that is when a mapped pointer must see the frame.
 */
#include "precompiled.h"
#include "commands.h"
#include "gpu.h"

#include "../include/visor_b200.h"

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

namespace
{
// Replays one command buffer. The stream is a packed sequence of {uint16 id}{payload} with unaligned
// payloads (cmd_alloc.cpp:59-70); every payload type gets an apply() overload, and run() dispatches on the
// id. The tracked GPUState does not persist across command buffers (cmd_exec.cpp:20).
class Replay
{
public:
  Replay(const byte *begin, const byte *end) : at(begin), stop(end) { memset(&gpu, 0, sizeof(gpu)); }

  void run()
  {
    while(at < stop)
    {
      Command id;
      fetch(id);
      switch(id)
      {
#define VB200_CMD(NAME)      \
  case Command::NAME:        \
  {                          \
    cmd::NAME payload;       \
    fetch(payload);          \
    apply(payload);          \
    break;                   \
  }
        VB200_CMD(PipelineBarrier)
        VB200_CMD(BeginRenderPass)
        VB200_CMD(EndRenderPass)
        VB200_CMD(BindPipeline)
        VB200_CMD(BindDescriptorSets)
        VB200_CMD(BindIB)
        VB200_CMD(BindVB)
        VB200_CMD(SetViewport)
        VB200_CMD(SetScissors)
        VB200_CMD(PushConstants)
        VB200_CMD(Draw)
        VB200_CMD(DrawIndexed)
        VB200_CMD(CopyBuf2Img)
        VB200_CMD(CopyBuf)
#undef VB200_CMD
      }
    }
  }

private:
  const byte *at, *stop;
  GPUState gpu;

  template <typename T>
  void fetch(T &v)
  {
    memcpy(&v, at, sizeof(T));
    at += sizeof(T);
  }

  // ---- commands without an effect on this path
  void apply(const cmd::PipelineBarrier &) {}
  void apply(const cmd::SetScissors &) {}
  void apply(const cmd::SetViewport &c) { gpu.view = c.view; }    // recorded, never read by the rasterizer

  // ---- render pass: subpass 0, colour attachment 0; clear values are consumed in the order of the
  // attachments that clear (cmd_exec.cpp:39-61)
  void apply(const cmd::BeginRenderPass &c)
  {
    const VkRenderPass_T::Subpass &sp = c.renderPass->subpasses[0];
    const VkClearValue *clear = c.clearval;
    gpu.col[0] = c.framebuffer->attachments[sp.colAttachments[0].idx]->image;
    if(sp.colAttachments[0].clear)
      ClearTarget(gpu.col[0], (clear++)->color);
    if(sp.depthAttachment.idx < 0)
      return;
    gpu.depth = c.framebuffer->attachments[sp.depthAttachment.idx]->image;
    if(sp.depthAttachment.clear)
      ClearTarget(gpu.depth, (clear++)->depthStencil);
  }
  void apply(const cmd::EndRenderPass &) { gpu.col[0] = VK_NULL_HANDLE; }

  // ---- bindings
  void apply(const cmd::BindPipeline &c) { gpu.pipeline = c.pipeline; }
  void apply(const cmd::BindDescriptorSets &c) { gpu.sets[c.idx] = c.set; }
  void apply(const cmd::BindIB &c)
  {
    gpu.ib.indexType = c.indexType;
    gpu.ib.offset = c.offset;
    gpu.ib.buffer = c.buffer;
  }
  void apply(const cmd::BindVB &c)
  {
    gpu.vbs[c.slot].offset = c.offset;
    gpu.vbs[c.slot].buffer = c.buffer;
  }
  void apply(const cmd::PushConstants &c) { memcpy(gpu.pushconsts + c.offset, c.values, c.size); }

  // ---- draws: instance counts / vertexOffset are ignored by the reference (cmd_exec.cpp:129-142)
  void apply(const cmd::Draw &c) { DrawTriangles(gpu, c.vertexCount, c.firstVertex, false); }
  void apply(const cmd::DrawIndexed &c) { DrawTriangles(gpu, c.indexCount, c.firstIndex, true); }

  // ---- copies (cmd_exec.cpp:143-182: plain memcpys there): performed between the HBM mirrors, in stream
  // order with the draws. Only whole, tightly packed mips of one layer exist.
  void apply(const cmd::CopyBuf2Img &c)
  {
    vb200_buffer from = {c.srcBuffer->bytes, c.srcBuffer->size};
    vb200_image to;
    toImage(c.dstImage, to);
    const VkImageSubresourceLayers &sub = c.region.imageSubresource;
    if(vb200_copy_buffer_to_image(&from, c.region.bufferOffset, &to, sub.mipLevel, sub.baseArrayLayer) != VB200_OK)
      printf("vkCmdCopyBufferToImage: %s\n", vb200_last_error());
  }
  void apply(const cmd::CopyBuf &c)
  {
    vb200_buffer from = {c.srcBuffer->bytes, c.srcBuffer->size}, to = {c.dstBuffer->bytes, c.dstBuffer->size};
    if(vb200_copy_buffer(&from, c.region.srcOffset, &to, c.region.dstOffset, c.region.size) != VB200_OK)
      printf("vkCmdCopyBuffer: %s\n", vb200_last_error());
  }
};
}    // namespace

void VkCommandBuffer_T::execute() const
{
  Replay(commandStream.data(), commandStream.data() + commandStream.size()).run();
}

// The reference's submit is synchronous and its fences / semaphores / WaitIdle are no-ops
// (cmd_exec.cpp:187-201, icd_stubs.cpp:64-155): everything must be host-visible when this returns.
VKAPI_ATTR VkResult VKAPI_CALL vkQueueSubmit(VkQueue, uint32_t submitCount, const VkSubmitInfo *submits, VkFence)
{
  for(const VkSubmitInfo *s = submits; s != submits + submitCount; s++)
    for(uint32_t i = 0; i < s->commandBufferCount; i++)
      s->pCommandBuffers[i]->execute();
  return vb200_flush() == VB200_OK ? VK_SUCCESS : VK_ERROR_DEVICE_LOST;
}
