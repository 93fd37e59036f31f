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
// the stream is a packed sequence of {uint16 id}{payload}; payloads are not aligned (cmd_alloc.cpp:59-70)
template <typename T>
T take(const byte *&cur)
{
  T v;
  memcpy(&v, cur, sizeof(T));
  cur += sizeof(T);
  return v;
}
}    // namespace

void VkCommandBuffer_T::execute() const
{
  const byte *cur = commandStream.data();
  const byte *end = cur + commandStream.size();

  GPUState state;    // does not persist across command buffers (cmd_exec.cpp:20)
  memset(&state, 0, sizeof(state));

  while(cur < end)
  {
    switch(take<Command>(cur))
    {
      case Command::PipelineBarrier: cur += sizeof(cmd::PipelineBarrier); break;
      case Command::BeginRenderPass:
      {
        // subpass 0, colour attachment 0; clear values consumed in cleared-attachment order (:39-61)
        const cmd::BeginRenderPass d = take<cmd::BeginRenderPass>(cur);
        const VkRenderPass_T::Subpass &sub = d.renderPass->subpasses[0];
        int clearIdx = 0;
        state.col[0] = d.framebuffer->attachments[sub.colAttachments[0].idx]->image;
        if(sub.colAttachments[0].clear)
          ClearTarget(state.col[0], d.clearval[clearIdx++].color);
        if(sub.depthAttachment.idx >= 0)
        {
          state.depth = d.framebuffer->attachments[sub.depthAttachment.idx]->image;
          if(sub.depthAttachment.clear)
            ClearTarget(state.depth, d.clearval[clearIdx++].depthStencil);
        }
        break;
      }
      case Command::EndRenderPass:
        cur += sizeof(cmd::EndRenderPass);
        state.col[0] = VK_NULL_HANDLE;
        break;
      case Command::BindPipeline: state.pipeline = take<cmd::BindPipeline>(cur).pipeline; break;
      case Command::BindDescriptorSets:
      {
        const cmd::BindDescriptorSets d = take<cmd::BindDescriptorSets>(cur);
        state.sets[d.idx] = d.set;
        break;
      }
      case Command::BindIB:
      {
        const cmd::BindIB d = take<cmd::BindIB>(cur);
        state.ib.buffer = d.buffer;
        state.ib.offset = d.offset;
        state.ib.indexType = d.indexType;
        break;
      }
      case Command::BindVB:
      {
        const cmd::BindVB d = take<cmd::BindVB>(cur);
        state.vbs[d.slot].buffer = d.buffer;
        state.vbs[d.slot].offset = d.offset;
        break;
      }
      case Command::SetViewport: state.view = take<cmd::SetViewport>(cur).view; break;    // recorded, never read
      case Command::SetScissors: cur += sizeof(cmd::SetScissors); break;
      case Command::PushConstants:
      {
        const cmd::PushConstants d = take<cmd::PushConstants>(cur);
        memcpy(state.pushconsts + d.offset, d.values, d.size);
        break;
      }
      case Command::Draw:
      {
        // instanceCount / firstInstance ignored (:129-135)
        const cmd::Draw d = take<cmd::Draw>(cur);
        DrawTriangles(state, d.vertexCount, d.firstVertex, false);
        break;
      }
      case Command::DrawIndexed:
      {
        // vertexOffset / instanceCount ignored (:136-142)
        const cmd::DrawIndexed d = take<cmd::DrawIndexed>(cur);
        DrawTriangles(state, d.indexCount, d.firstIndex, true);
        break;
      }
      case Command::CopyBuf2Img:
      {
        // whole tightly packed mip of one layer only (:147-164); performed between the HBM mirrors, in
        // stream order with the draws
        const cmd::CopyBuf2Img d = take<cmd::CopyBuf2Img>(cur);
        vb200_buffer src = {d.srcBuffer->bytes, d.srcBuffer->size};
        vb200_image dst;
        toImage(d.dstImage, dst);
        if(vb200_copy_buffer_to_image(&src, d.region.bufferOffset, &dst, d.region.imageSubresource.mipLevel,
                                      d.region.imageSubresource.baseArrayLayer) != VB200_OK)
          printf("vkCmdCopyBufferToImage: %s\n", vb200_last_error());
        break;
      }
      case Command::CopyBuf:
      {
        const cmd::CopyBuf d = take<cmd::CopyBuf>(cur);
        vb200_buffer src = {d.srcBuffer->bytes, d.srcBuffer->size}, dst = {d.dstBuffer->bytes, d.dstBuffer->size};
        if(vb200_copy_buffer(&src, d.region.srcOffset, &dst, d.region.dstOffset, d.region.size) != VB200_OK)
          printf("vkCmdCopyBuffer: %s\n", vb200_last_error());
        break;
      }
    }
  }
}

VKAPI_ATTR VkResult VKAPI_CALL vkQueueSubmit(VkQueue queue, uint32_t submitCount, const VkSubmitInfo *pSubmits,
                                             VkFence fence)
{
  for(uint32_t i = 0; i < submitCount; i++)
    for(uint32_t c = 0; c < pSubmits[i].commandBufferCount; c++)
      pSubmits[i].pCommandBuffers[c]->execute();
  // fences/semaphores/WaitIdle are no-ops in the reference (icd_stubs.cpp:64-155): everything must be
  // host-visible when this returns
  return vb200_flush() == VB200_OK ? VK_SUCCESS : VK_ERROR_DEVICE_LOST;
}
