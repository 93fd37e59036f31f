// runtime_internal.h — what runtime.cpp and mgpu.cpp share (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <vector>

namespace vb200
{
// error reporting of the C-ABI (vb200_last_error)
int set_error(int code, const char *fmt, ...);
// the library stream; NULL before vb200_init
cudaStream_t library_stream();
int library_device();
void count_launches(int n);

// Fused sort-first exchange targets: [local, local + bytes) is a colour image (or a buffer holding one) on this
// rank; peers[i] is the same range on the other ranks, mapped into this GPU's address space; multicast (may
// be NULL) is an NVSwitch multicast mapping of the range on all ranks. The tile kernels store every colour
// word they produce at local + offset, and at the same offset of the multicast mapping (or of each peer).
void set_exchange_range(uint8_t *local, size_t bytes, const std::vector<uint8_t *> &peers, uint8_t *multicast);
void clear_exchange_range(uint8_t *local);

// multi-GPU state kept by mgpu.cpp
bool mgpu_active();
// symmetric allocation for HBM mirrors of registered host ranges (option "mgpu_mirrors"); collective.
int mgpu_sym_alloc(size_t bytes, uint8_t **local);
int mgpu_sym_free(uint8_t *local);
bool mgpu_is_symmetric(const uint8_t *dev);
}    // namespace vb200
