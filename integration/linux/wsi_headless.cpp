/*
 * integration/linux/wsi_headless.cpp — headless stand-in for wsi.cpp (Win32 GDI swapchain, not buildable
 * on Linux). icd_interface.cpp:23-26,41-45 takes the address of these entry points; rendering tests use
 * plain VkImages bound to host-visible memory, never a swapchain. Part of the Linux build of the CUDA ICD
 * (integration/libvisor_b200_icd.so); the checker's build of the reference (oracle/_ref/libvisor_ref.so)
 * compiles it too. Contains no reference code.
 */
#include "precompiled.h"

VKAPI_ATTR VkResult VKAPI_CALL vkGetPhysicalDeviceSurfaceSupportKHR(VkPhysicalDevice, uint32_t,
                                                                    VkSurfaceKHR, VkBool32 *pSupported)
{
  *pSupported = VK_FALSE;
  return VK_SUCCESS;
}
VKAPI_ATTR VkResult VKAPI_CALL vkGetPhysicalDeviceSurfaceFormatsKHR(VkPhysicalDevice, VkSurfaceKHR,
                                                                    uint32_t *pCount,
                                                                    VkSurfaceFormatKHR *)
{
  *pCount = 0;
  return VK_SUCCESS;
}
VKAPI_ATTR VkResult VKAPI_CALL vkGetPhysicalDeviceSurfaceCapabilitiesKHR(VkPhysicalDevice, VkSurfaceKHR,
                                                                         VkSurfaceCapabilitiesKHR *p)
{
  memset(p, 0, sizeof(*p));
  return VK_ERROR_SURFACE_LOST_KHR;
}
VKAPI_ATTR VkResult VKAPI_CALL vkGetPhysicalDeviceSurfacePresentModesKHR(VkPhysicalDevice, VkSurfaceKHR,
                                                                         uint32_t *pCount,
                                                                         VkPresentModeKHR *)
{
  *pCount = 0;
  return VK_SUCCESS;
}
VKAPI_ATTR VkResult VKAPI_CALL vkCreateSwapchainKHR(VkDevice, const VkSwapchainCreateInfoKHR *,
                                                    const VkAllocationCallbacks *, VkSwapchainKHR *)
{
  return VK_ERROR_SURFACE_LOST_KHR;
}
VKAPI_ATTR void VKAPI_CALL vkDestroySwapchainKHR(VkDevice, VkSwapchainKHR, const VkAllocationCallbacks *)
{
}
VKAPI_ATTR VkResult VKAPI_CALL vkGetSwapchainImagesKHR(VkDevice, VkSwapchainKHR, uint32_t *pCount,
                                                       VkImage *)
{
  *pCount = 0;
  return VK_SUCCESS;
}
VKAPI_ATTR VkResult VKAPI_CALL vkAcquireNextImageKHR(VkDevice, VkSwapchainKHR, uint64_t, VkSemaphore,
                                                     VkFence, uint32_t *)
{
  return VK_ERROR_SURFACE_LOST_KHR;
}
VKAPI_ATTR VkResult VKAPI_CALL vkQueuePresentKHR(VkQueue, const VkPresentInfoKHR *)
{
  return VK_ERROR_SURFACE_LOST_KHR;
}

