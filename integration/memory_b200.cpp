/*
 * integration/memory_b200.cpp — VkDeviceMemory for the CUDA path; link INSTEAD OF memory.cpp.
 *
 * Semantics are the reference's (memory.cpp:5-41): memory is host memory, vkMapMemory returns
 * bytes + offset, flushes are no-ops, buffers/images alias it at bind time. The only difference is
 * that the allocation is registered with the CUDA library, which page-locks it and creates its HBM
 * mirror, so the per-submit uploads/downloads run at full PCIe speed; DEVICE_LOCAL allocations (memory
 * type 0) live in the mirror only.
 */
#include "precompiled.h"

#include "../include/visor_b200.h"

namespace
{
// texel fetches of 1-byte formats read 4 bytes per texel, block loads 16: keep a tail past every allocation
const size_t kTail = 16;
// page-aligned host storage: page-locking (cudaHostRegister inside vb200_mem_register) then covers exactly
// this allocation's pages
const size_t kAlign = 4096;

byte *hostStorage(VkDeviceSize size, bool deviceLocal)
{
  const size_t bytes = ((size_t)size + kTail + kAlign - 1) / kAlign * kAlign;
  byte *p = (byte *)aligned_alloc(kAlign, bytes);
  if(!p)
    return NULL;
  memset(p, 0, bytes);
  if(vb200_mem_register(p, size + kTail) != VB200_OK)
    printf("visor_b200: vkAllocateMemory: %s\n", vb200_last_error());
  // memory type 0 is DEVICE_LOCAL only (query.cpp:246-267): the application cannot map it, so the HBM
  // mirror is the resource; it is filled by vkCmdCopyBuffer* on the device and never crosses PCIe again
  else if(deviceLocal)
    vb200_mem_set_device_local(p, 1);
  return p;
}

void releaseStorage(byte *p)
{
  if(!p)
    return;
  vb200_mem_unregister(p);
  free(p);
}
}    // namespace

VKAPI_ATTR VkResult VKAPI_CALL vkAllocateMemory(VkDevice, const VkMemoryAllocateInfo *info, const VkAllocationCallbacks *,
                                                VkDeviceMemory *out)
{
  byte *storage = hostStorage(info->allocationSize, info->memoryTypeIndex == 0);
  if(!storage)
    return VK_ERROR_OUT_OF_HOST_MEMORY;
  *out = new VkDeviceMemory_T;
  (*out)->size = info->allocationSize;
  (*out)->bytes = storage;
  return VK_SUCCESS;
}

VKAPI_ATTR void VKAPI_CALL vkFreeMemory(VkDevice, VkDeviceMemory mem, const VkAllocationCallbacks *)
{
  if(mem)
  {
    releaseStorage(mem->bytes);
    delete mem;
  }
}

// memory.cpp:22-41: a mapping is the host pointer itself; flushes are no-ops because the library re-reads
// host memory at the first use after every submit (coherent mode)
VKAPI_ATTR VkResult VKAPI_CALL vkMapMemory(VkDevice, VkDeviceMemory mem, VkDeviceSize offset, VkDeviceSize, VkMemoryMapFlags,
                                           void **mapped)
{
  *mapped = mem->bytes + offset;
  return VK_SUCCESS;
}

VKAPI_ATTR void VKAPI_CALL vkUnmapMemory(VkDevice, VkDeviceMemory)
{
}

VKAPI_ATTR VkResult VKAPI_CALL vkFlushMappedMemoryRanges(VkDevice, uint32_t, const VkMappedMemoryRange *)
{
  return VK_SUCCESS;
}
