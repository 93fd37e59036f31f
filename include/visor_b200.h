/*
 * visor_b200.h — C-ABI of the B200-native draw-execution path for visor.
 *
 * This is the drop-in boundary.  Every entry point replaces one function of
 * visor's internal "GPU" operator API; the reference declaration it stands in
 * for is cited next to it (paths relative to the reference tree).  The ABI is
 * plain C: pointers, sizes and PODs only.  All `void*` data pointers that
 * cross it are HOST pointers exactly as in the reference (VkBuffer_T::bytes,
 * VkImage_T::pixels — precompiled.h:89-109) unless they are CUDA device
 * pointers, which the library detects (cudaPointerGetAttributes) and uses in
 * place.  Host ranges are mirrored in HBM by the library (see vb200_mem_*).
 *
 * Enum-valued fields carry the reference's own Vulkan enum values
 * (3rdparty/vulkan.h, header version 42) so the binding passes them through
 * untouched.
 *
 * Threading: like the reference (DrawTriangles is not re-entrant,
 * rasterizer.cpp:370,374), one thread at a time per process.
 * Errors: every call returns 0 on success or a negative vb200_status; the
 * message of the last failure is kept and returned by vb200_last_error().
 * Nothing throws across this boundary.  There is no CPU fallback: without a
 * usable CUDA device every compute entry point fails with VB200_ERR_NO_DEVICE.
 */
#ifndef VISOR_B200_H
#define VISOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_ABI_VERSION 1

#if defined(__GNUC__)
#define VB200_API __attribute__((visibility("default")))
#else
#define VB200_API
#endif

typedef enum vb200_status {
  VB200_OK = 0,
  VB200_ERR_NO_DEVICE = -1,    /* no CUDA device / driver; nothing was computed */
  VB200_ERR_CUDA = -2,         /* a CUDA runtime call failed (sticky)            */
  VB200_ERR_SPIRV = -3,        /* module outside the reference's SPIR-V subset   */
  VB200_ERR_LINK = -4,         /* ptxas / module load failed                     */
  VB200_ERR_INVALID = -5,      /* bad argument / unsupported state               */
  VB200_ERR_NOT_INITIALIZED = -6
} vb200_status;

/* ---- resources: mirror VkImage_T / VkBuffer_T (precompiled.h:89-109) -------------------- */

typedef struct vb200_image {
  void *pixels;            /* VkImage_T::pixels (host) or a device pointer */
  uint32_t width, height, depth; /* VkImage_T::extent */
  uint32_t image_type;     /* VkImageType   (not read by the path)       */
  uint32_t format;         /* VkFormat                                    */
  uint32_t array_layers;
  uint32_t mip_levels;
  uint32_t bytes_per_pixel;
} vb200_image;

typedef struct vb200_buffer {
  void *bytes;   /* VkBuffer_T::bytes (host) or a device pointer */
  uint64_t size; /* VkBuffer_T::size                              */
} vb200_buffer;

/* ---- shaders: replace spirv_compile.h:3-10 ---------------------------------------------- */

typedef struct vb200_shader vb200_shader; /* stands in for LLVMFunction (opaque, precompiled.h:49) */
typedef struct vb200_entry vb200_entry;   /* stands in for the Shader function pointer (precompiled.h:52-55) */

/* InitLLVM / ShutdownLLVM (spirv_compile.h:3-4) and InitRasterThreads / ShutdownRasterThreads
 * (gpu.h:69-70) collapse into one init/shutdown pair: create the CUDA context + stream on
 * `device`, load the kernel scaffolds.  vb200_init is idempotent. */
VB200_API int vb200_init(int device);
VB200_API void vb200_shutdown(void);

/* CompileFunction(const uint32_t *pCode, size_t codeSize) — spirv_compile.h:8, spirv_compile.cpp:645.
 * Parses the module and lowers every entry point to a PTX device function.  Returns NULL on failure
 * (the reference's NULL -> VK_ERROR_DEVICE_LOST convention, shaders.cpp:13-14). Needs no device. */
VB200_API vb200_shader *vb200_shader_create(const uint32_t *code, size_t code_size_words);
/* GetFuncPointer(LLVMFunction*, const char *name) — spirv_compile.h:9, spirv_compile.cpp:2434. */
VB200_API vb200_entry *vb200_shader_entry(vb200_shader *shader, const char *name);
/* DestroyFunction(LLVMFunction*) — spirv_compile.h:10. Entries of the module die with it. */
VB200_API void vb200_shader_destroy(vb200_shader *shader);
/* Debug/inspection: PTX text generated for an entry point (owned by the module). */
VB200_API const char *vb200_entry_ptx(const vb200_entry *entry);
/* Debug/CI: run only the compile step for a VS/FS pair (each kernel's PTX + the shader function it calls ->
 * one sm_100a cubin per kernel, through ptxas). Needs no device; reports the total cubin size. */
VB200_API int vb200_link_check(const vb200_entry *vs, const vb200_entry *fs, uint64_t *cubin_size);
/* 0 = vertex, 4 = fragment (spv::ExecutionModel). */
VB200_API int vb200_entry_stage(const vb200_entry *entry);
/* The descriptors an entry point dereferences, so the binding can pick exactly those out of
 * GPUState::sets[] (VkDescriptorSet_T does not record how many binds it holds, precompiled.h:140-156). */
VB200_API int vb200_entry_num_resources(const vb200_entry *entry);
VB200_API int vb200_entry_resource(const vb200_entry *entry, int index, uint32_t *set, uint32_t *binding,
                                   uint32_t *is_image);

/* ---- fixed-function + draw state: mirror VkPipeline_T / GPUState ------------------------ */

typedef struct vb200_vertex_attr { /* VkPipeline_T::vattrs[] (precompiled.h:118-124) */
  uint32_t format;                 /* VkFormat; 0 (UNDEFINED) = attribute unused */
  uint32_t stride;
  uint32_t offset;
  uint32_t vb;
} vb200_vertex_attr;

typedef struct vb200_pipeline { /* VkPipeline_T (precompiled.h:116-133) */
  vb200_vertex_attr vattrs[16];
  uint32_t topology;           /* VkPrimitiveTopology: TRIANGLE_LIST(3) / TRIANGLE_STRIP(4)     */
  uint32_t front_face;         /* VkFrontFace                                                   */
  uint32_t cull_mode;          /* VkCullModeFlags                                               */
  uint32_t depth_compare_op;   /* VkCompareOp; ALWAYS(7) = no test (rasterizer.cpp:562)         */
  uint32_t depth_write_enable; /* bool                                                          */
  uint32_t blend_enable;       /* VkPipelineColorBlendAttachmentState::blendEnable              */
  uint32_t src_color_blend_factor;
  uint32_t dst_color_blend_factor;
  uint32_t color_blend_op;
  const vb200_entry *vs; /* VkPipeline_T::vs */
  const vb200_entry *fs; /* VkPipeline_T::fs */
} vb200_pipeline;

/* One descriptor of a bound set: VkDescriptorSet_T::Bind (precompiled.h:140-156) flattened
 * with its (set, binding) coordinates. */
typedef struct vb200_binding {
  uint32_t set;
  uint32_t binding;
  uint32_t type;       /* VkDescriptorType (informational)                                   */
  uint32_t is_image;   /* 0: buffer/offset valid (bufferInfo); 1: image valid (imageInfo)    */
  vb200_buffer buffer; /* VkDescriptorBufferInfo::buffer                                     */
  uint64_t offset;     /* VkDescriptorBufferInfo::offset                                     */
  vb200_image image;   /* imageInfo.imageView->image                                         */
} vb200_binding;

typedef struct vb200_draw_state { /* GPUState (gpu.h:3-23) */
  struct {
    vb200_buffer buffer;
    uint64_t offset;
    uint32_t index_type; /* VkIndexType: UINT16(0) / UINT32(1) */
    uint32_t _pad;
  } ib;
  struct {
    vb200_buffer buffer;
    uint64_t offset;
  } vbs[4];
  vb200_image color; /* GPUState::col[0]                         */
  vb200_image depth; /* GPUState::depth; pixels == NULL -> none  */
  const vb200_pipeline *pipeline;
  const vb200_binding *bindings; /* all binds of GPUState::sets[] the shaders may touch */
  uint32_t num_bindings;
  uint32_t _pad;
  uint8_t pushconsts[128];
} vb200_draw_state;

/* ---- the operators ---------------------------------------------------------------------- */

/* void ClearTarget(VkImage, const VkClearColorValue&) — gpu.h:59, rasterizer.cpp:332-361. */
VB200_API int vb200_clear_color(const vb200_image *target, const float rgba[4]);
/* void ClearTarget(VkImage, const VkClearDepthStencilValue&) — gpu.h:60, rasterizer.cpp:312-330. */
VB200_API int vb200_clear_depth(const vb200_image *target, float depth);
/* void DrawTriangles(const GPUState&, int numVerts, uint32_t first, bool indexed) — gpu.h:61,
 * rasterizer.cpp:363-520.  Asynchronous: results are host-visible after vb200_flush(). */
VB200_API int vb200_draw(const vb200_draw_state *state, int num_verts, uint32_t first, int indexed);
/* Stand-alone texture unit, for parity tests of the sampler:
 * sample_tex_wrapped(u, v, tex, byteOffs, out) — gpu.h:64-65, texture_sampling.cpp:139-184 and
 * sample_cube_wrapped(x, y, z, tex, out) — gpu.h:66-67, texture_sampling.cpp:186-250, evaluated on
 * the device for `count` coordinates (host arrays; uvw has 2 or 3 floats per sample). */
VB200_API int vb200_sample(const vb200_image *tex, int cube, uint64_t byte_offset, const float *uvw,
                           float *out_rgba, size_t count);
/* End of a submit (the reference's vkQueueSubmit is synchronous, cmd_exec.cpp:187-201): waits for
 * the stream and, in coherent mode, copies every attachment written since the last flush back to
 * its host range so a mapped pointer sees it. */
VB200_API int vb200_flush(void);

/* ---- residency: VkDeviceMemory semantics (memory.cpp:5-41) ------------------------------- */

typedef enum vb200_sync_mode {
  /* Default, drop-in: host memory is authoritative between submits. Every host range a draw reads
   * is re-uploaded at its first use after a flush; attachments are downloaded by vb200_flush. */
  VB200_SYNC_COHERENT = 0,
  /* Caller moves data explicitly with vb200_mem_upload/_download (HBM-resident resources,
   * memory type 0 "DEVICE_LOCAL", query.cpp:260-266). Draws never copy. */
  VB200_SYNC_EXPLICIT = 1
} vb200_sync_mode;

VB200_API int vb200_set_sync_mode(int mode);
/* Create (or find) the HBM mirror of [host, host+size) and page-lock the host range. Draws
 * auto-register ranges they have not seen. */
VB200_API int vb200_mem_register(void *host, uint64_t size);
VB200_API int vb200_mem_unregister(void *host);
VB200_API int vb200_mem_upload(const void *host, uint64_t size);   /* host -> HBM mirror, async */
VB200_API int vb200_mem_download(void *host, uint64_t size);       /* HBM mirror -> host, async */
/* The host is about to write [host, host+size) mid-submit (vkCmdCopyBuffer / vkCmdCopyBufferToImage are
 * replayed as memcpy, cmd_exec.cpp:143-182): orders the write after in-flight uploads of the range and
 * makes later draws re-upload it. */
VB200_API int vb200_mem_host_write(const void *host, uint64_t size);
/* Marks a registered range as DEVICE_LOCAL memory (memory type 0, query.cpp:260-266): the application
 * never maps it, so its HBM mirror is the only copy that matters. It is filled by vb200_copy_* (or
 * vb200_mem_upload), never re-uploaded per submit and never downloaded by vb200_flush. */
VB200_API int vb200_mem_set_device_local(void *host, int device_local);
/* vkCmdCopyBuffer as the reference replays it (cmd_exec.cpp:176-182): memcpy(dst->bytes + dst_offset,
 * src->bytes + src_offset, size) — performed between the HBM mirrors on the library stream, in
 * submission order with the draws. In coherent mode the destination is downloaded by vb200_flush like
 * an attachment, so host memory ends up identical to the reference's. */
VB200_API int vb200_copy_buffer(const vb200_buffer *src, uint64_t src_offset, const vb200_buffer *dst,
                                uint64_t dst_offset, uint64_t size);
/* vkCmdCopyBufferToImage as the reference replays it (cmd_exec.cpp:143-174): one whole, tightly packed
 * mip level of one array layer; destination offset = CalcSubresourceByteOffset(dst, mip, layer)
 * (precompiled.cpp:3-36), byte count = max(1,w>>mip) * max(1,h>>mip) * bytes_per_pixel. */
VB200_API int vb200_copy_buffer_to_image(const vb200_buffer *src, uint64_t buffer_offset, const vb200_image *dst,
                                         uint32_t mip_level, uint32_t array_layer);
/* Present / read-back path (the reference's vkQueuePresentKHR blits the finished swapchain image to the
 * window, wsi.cpp:86-219; headless, the "window" is host memory). Copies `image` as it is after all work
 * queued so far to `dst_host` on a SECOND stream and returns at once: the library stream goes on with
 * the next frame while the copy crosses PCIe. Work that later WRITES the image is ordered after the
 * copy automatically, so a double-buffered application overlaps fully and a single-buffered one stays
 * correct. `dst_host` should be page-locked (inside a vb200_mem_register'ed range) for a truly
 * asynchronous copy. vb200_present_wait blocks until the copy of that ticket has landed;
 * vb200_flush waits for all of them. */
VB200_API int vb200_present(const vb200_image *image, void *dst_host, uint64_t dst_size, int *ticket);
VB200_API int vb200_present_wait(int ticket);
/* Device address of a mirrored host pointer (NULL if not mirrored); for interop (NCCL, torch). */
VB200_API void *vb200_mem_device_ptr(const void *host);

/* ---- sort-first multi-GPU ---------------------------------------------------------------- */

/* Restrict rasterisation to screen tiles t with t % world == rank (tile index = ty*tiles_x+tx,
 * 32x32 px). Geometry stages run on all ranks. rank 0 / world 1 restores single-GPU behaviour. */
VB200_API int vb200_set_tile_owner(int rank, int world);
/* Pack the owned colour tiles of `image` into `dst` (device pointer, owner-major: the k-th owned
 * tile occupies 4096 bytes at dst + k*4096), or scatter a gathered buffer (world*slots_per_rank
 * tiles, rank-major) back into the linear image. Both run on the library stream. */
VB200_API int vb200_tiles_pack(const vb200_image *image, void *dst_device, uint64_t dst_size);
VB200_API int vb200_tiles_unpack(const vb200_image *image, const void *src_device, uint64_t src_size,
                                 int world);
VB200_API uint32_t vb200_tiles_per_rank(uint32_t width, uint32_t height, int world);
/* Fused exchange (preferred over pack/all-gather/unpack): `local_color_device` is this rank's colour
 * image (a device pointer used as vb200_draw_state::color.pixels); `peer_color_device[i]` are the
 * peer-mapped device addresses of the SAME image on the other ranks (e.g. torch symmetric memory
 * buffer_ptrs, CUDA IPC or cuMem handles). While tile ownership is on, the tile kernels store every
 * colour word they produce into all of them over NVLink, so after all ranks finish (one cross-rank
 * barrier) every rank holds the whole image. num_peers = 0 removes the association. */
VB200_API int vb200_set_peer_targets(const void *local_color_device, void *const *peer_color_device, int num_peers);
/* Same exchange through the NVSwitch: `multicast_device` is a multicast (NVLS) mapping that spans the
 * image on ALL ranks (e.g. torch symmetric memory multicast_ptr). Each colour word is then sent once
 * (multimem.st) and replicated to every rank by the switch, instead of once per peer. Takes
 * precedence over peer targets; NULL removes the association. */
VB200_API int vb200_set_multicast_target(const void *local_color_device, void *multicast_device);

/* ---- sort-first over the GPUs of one node, one process per GPU (no framework in the data plane) ----
 * vb200_mgpu_init joins the calling process to `session` (any string shared by the `world` <= 8 processes,
 * e.g. the launcher's port number) as `rank`, running on CUDA device `device`: it bootstraps over abstract
 * unix-domain sockets, calls vb200_set_tile_owner(rank, world) and sets up the device-side barrier. All
 * vb200_mgpu_* calls below except _info and _push are COLLECTIVE: every rank makes them in the same order.
 * vb200_mgpu_alloc returns this rank's copy of a symmetric buffer (one allocation per rank, all of them
 * mapped into every rank; where the NVSwitch supports it also one multicast mapping). A colour target that
 * lies in a symmetric buffer gets the fused exchange: the tile kernels store every pixel into all ranks'
 * copies (one multimem.st, or one store per peer), so after vb200_mgpu_barrier every rank holds the image.
 * vb200_mgpu_barrier enqueues a cross-rank barrier on the library stream (it gives up after ~2 s if a
 * rank never arrives; the next call then reports it).
 * vb200_mgpu_push replicates bytes [p, p + bytes) of this rank's copy of a symmetric buffer to the same
 * range of every other rank (offset and size multiples of 16).
 * With option "mgpu_mirrors" = 1 the HBM mirrors of ranges registered afterwards (vb200_mem_register, a
 * collective then) are symmetric buffers: colour targets in host memory get the fused exchange,
 * vb200_flush ends with the barrier, and vb200_mgpu_upload(host, size) uploads a frame input in slices —
 * each rank sends 1/world of it over its own PCIe link and pushes that slice to the others over NVLink
 * (follow the uploads of a frame with one vb200_mgpu_barrier before drawing). Depth is not exchanged. */
VB200_API int vb200_mgpu_init(int rank, int world, int device, const char *session);
VB200_API int vb200_mgpu_shutdown(void);
VB200_API int vb200_mgpu_info(int *rank, int *world, int *multicast);
VB200_API int vb200_mgpu_alloc(uint64_t bytes, void **local_device);
VB200_API int vb200_mgpu_free(void *local_device);
VB200_API int vb200_mgpu_barrier(void);
VB200_API int vb200_mgpu_push(const void *local_device, uint64_t bytes);
VB200_API int vb200_mgpu_upload(const void *host, uint64_t size);

/* ---- introspection ----------------------------------------------------------------------- */

typedef struct vb200_stats { /* counters of the reference (rasterizer.cpp:517-519,693-695) + timing */
  uint64_t draws;
  uint64_t triangles_in;
  uint64_t triangles_out;      /* survived degenerate + cull                      */
  uint64_t tile_pairs;         /* (triangle, tile) list entries produced by binning */
  uint64_t fragments_covered;  /* "pixels/written" of the reference                */
  uint64_t fragments_shaded;   /* "depth/passed"                                   */
  uint64_t kernel_launches;    /* launches of this library's kernels               */
  uint64_t h2d_bytes, d2h_bytes;
} vb200_stats;

/* Benchmark hygiene: writes a 512 MB scratch buffer on the library stream (evicts the 126 MB L2). */
VB200_API int vb200_l2_flush(void);
/* Device-side timing for benchmarks: CUDA events recorded on the library stream (slots 0..15). */
VB200_API int vb200_event_record(int slot);
VB200_API int vb200_event_elapsed_ms(int start_slot, int end_slot, float *ms); /* syncs on end_slot */
/* Per-phase device time accumulated while option "time_kernels" is 1 (diagnostic: it syncs per draw). */
enum { VB200_PHASE_CLEAR = 0, VB200_PHASE_VERTEX = 1, VB200_PHASE_SETUP = 2, VB200_PHASE_BIN = 3,
       VB200_PHASE_TILES = 4, VB200_PHASES = 5 };
VB200_API int vb200_get_phase_times(double *ms, uint64_t *counts, int n);
VB200_API int vb200_get_stats(vb200_stats *out);   /* syncs the stream */
VB200_API void vb200_reset_stats(void);
VB200_API void *vb200_stream(void);                /* cudaStream_t the library launches on */
VB200_API const char *vb200_last_error(void);
VB200_API int vb200_abi_version(void);
/* Name of the tile kernel the most recent vb200_draw launched ("vb200_k_tile_ordered",
 * "vb200_k_tile_resolve_min_first", ...; "" before the first draw). For benchmark reports. */
VB200_API const char *vb200_last_tile_kernel(void);
/* Tuning/debug knobs by name ("raster_path": 0 auto, 1 ordered tiles; "count_fragments": 0/1,
 * "time_kernels": 0/1, "fuse_clears": 0/1, "slot_keys": 0/1 — 0 forces the resolve kernels' code path of
 * draws with 2^24 or more triangles; "tile_list_cap": entries per tile list, 0 = automatic — a tiny value
 * forces the tile kernels' fallback for overflowed lists; "mgpu_mirrors": see vb200_mgpu_init).
 * "extended_spirv": 1 makes later vb200_shader_create calls accept opcodes the reference asserts on: OpPhi,
 * OpSwitch, OpKill (discard; such fragment shaders always run on the in-order tile kernel), OpSelect, the
 * remaining ordered and all unordered float comparisons, the remaining integer comparisons, integer division /
 * remainder / shifts / bit, bit-field and logic operations, OpBitcast, OpConvertFToS/FToU/UToF, OpIsNan/OpIsInf,
 * OpAny/OpAll, OpFRem/OpFMod, OpCopyObject, OpUndef, OpConstantNull/True/False, OpCompositeInsert,
 * OpVectorExtractDynamic/InsertDynamic, OpImageSampleExplicitLod (level ignored: mip 0) and 45 GLSL.std.450
 * instructions (DESIGN.md section 3 lists them, the results fixed where SPIR-V leaves them open, and the
 * fifteen transcendental ones that are approximate like the reference subset's Sin/Cos/Pow) — off by default,
 * because with it the front end no longer rejects exactly what CompileFunction rejects
 * (spirv_compile.cpp:1734,1888). Unknown names return VB200_ERR_INVALID. */
VB200_API int vb200_set_option(const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* VISOR_B200_H */
