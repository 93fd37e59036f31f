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

VKAPI_ATTR VkResult VKAPI_CALL vkAllocateMemory(VkDevice device, const VkMemoryAllocateInfo *pAllocateInfo,
                                                const VkAllocationCallbacks *pAllocator, VkDeviceMemory *pMemory)
{
  VkDeviceMemory mem = new VkDeviceMemory_T;
  mem->size = pAllocateInfo->allocationSize;
  mem->bytes = new byte[mem->size + 16];    // +16: texel fetches of 1-byte formats read 4 bytes
  vb200_mem_register(mem->bytes, mem->size + 16);
  // memory type 0 is DEVICE_LOCAL only (query.cpp:246-267): the application cannot map it, so the HBM
  // mirror is the resource; it is filled by vkCmdCopyBuffer* on the device and never crosses PCIe again
  if(pAllocateInfo->memoryTypeIndex == 0)
    vb200_mem_set_device_local(mem->bytes, 1);
  *pMemory = mem;
  return VK_SUCCESS;
}

VKAPI_ATTR void VKAPI_CALL vkFreeMemory(VkDevice device, VkDeviceMemory memory, const VkAllocationCallbacks *pAllocator)
{
  if(!memory)
    return;
  vb200_mem_unregister(memory->bytes);
  delete[] memory->bytes;
  delete memory;
}

VKAPI_ATTR VkResult VKAPI_CALL vkMapMemory(VkDevice device, VkDeviceMemory memory, VkDeviceSize offset,
                                           VkDeviceSize size, VkMemoryMapFlags flags, void **ppData)
{
  *ppData = memory->bytes + offset;
  return VK_SUCCESS;
}

VKAPI_ATTR void VKAPI_CALL vkUnmapMemory(VkDevice device, VkDeviceMemory memory)
{
}

VKAPI_ATTR VkResult VKAPI_CALL vkFlushMappedMemoryRanges(VkDevice device, uint32_t memoryRangeCount,
                                                         const VkMappedMemoryRange *pMemoryRanges)
{
  return VK_SUCCESS;    // coherent: the library re-reads host memory at the first use after every submit
}
