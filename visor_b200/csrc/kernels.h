// kernels.h — host-callable launchers of the shader-independent kernels (fixed.cu) and the parameter
// blocks of the JIT-linked kernels (scaffold.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "device_types.h"

// A batch = consecutive vkCmdDraw*s of one submit that share pipeline, bindings and attachments (the reference
// replays them one by one, cmd_exec.cpp:129-142). They are rasterised together: ONE vertex kernel over the
// batch's vertex spans, ONE setup/binning kernel over all its triangles (triangle id = tri_base + local index,
// so ids keep submission order across draws) and ONE tile pass.
struct Vb200VertexSpan    // a run of source vertices shaded into consecutive post-VS record slots
{
  uint32_t src_base;      // first source vertex (vertex index, or index value for indexed draws)
  uint32_t count;
  uint32_t slot_base;     // first post-VS record slot
  uint32_t pad;
};
struct Vb200BatchDraw
{
  const void *ib;         // device address of ib.buffer->bytes + ib.offset; NULL: non-indexed draw
  uint32_t index_type, first, num_tris, tri_base;
  uint32_t span;          // the vertex span that holds the draw's corners
  uint32_t pad;
};

struct Vb200SetupParams
{
  // the batch: draw0/span0 are used in place of the tables when those are NULL (a batch of one draw)
  const Vb200BatchDraw *draws;
  const Vb200VertexSpan *spans;
  uint32_t num_draws;
  Vb200BatchDraw draw0;
  Vb200VertexSpan span0;
  // an indexed draw out of a much larger vertex buffer, too big to measure on the host, is always alone in its
  // batch: device {minIndex, maxIndex} (k_index_range); its span is [min, max], clipped to span0.count records
  const uint32_t *range;
  uint32_t num_tris, topology;    // triangles of the whole batch
  uint32_t vertex_bound;    // indexed draws: vertices the bound buffers hold; an index at or above it kills the triangle
  const Vb200RasterVertex *rv;
  Vb200TriRecord *tri;
  uint32_t *tri_tiles;    // packed tile range per triangle (VB200_TILES_DEAD = dead): the tile kernels' fallback list
  // Binning is one pass: the setup kernel appends each surviving triangle straight into the list of every
  // tile its bbox touches. Tile t (owned slot t / owner_world) has room for list_cap ids at
  // list[slot * list_cap]; tile_count[t] counts EVERY append, so a count above list_cap says "this list is
  // incomplete" and the tile kernel then finds its triangles by scanning tri_tiles instead (exact, in order,
  // no host round trip and no retry).
  uint32_t *tile_count;
  uint32_t *list;
  uint32_t list_cap;
  Vb200DrawCounters *counters;
  uint32_t front_face, cull_mode;
  uint32_t width, height, tiles_x, tiles_y, owner_rank, owner_world;
};

// parameter blocks of scaffold.cu (must match the definitions there)
struct Vb200VertexParams
{
  const Vb200VertexSpan *spans;    // NULL: span0 alone
  uint32_t num_spans;
  Vb200VertexSpan span0;
  const uint32_t *range;           // device {minIndex, maxIndex}: span0 = [min, max] clipped to span0.count records
  uint32_t count;                  // records to shade (all spans)
  uint32_t vertex_bound;    // vertices the bound vertex buffers hold (0xffffffff: unbounded): nothing beyond is fetched
  Vb200RasterVertex *rv;
  float4 *interps;
  uint32_t nslots;
  uint32_t width, height;
  // the per-tile counters the setup kernel (next in the stream) accumulates into: zeroed here, which
  // saves a memset node between the two kernels
  uint32_t *tile_count;
  uint32_t tile_count_n;
};
struct Vb200TileParams
{
  const Vb200TriRecord *tri;
  const Vb200RasterVertex *rv;
  const uint32_t *list;         // per-tile lists: tile t's ids start at list[(t / owner_world) * list_cap]
  const uint32_t *tile_count;   // appends per tile; > list_cap: the list is incomplete, scan tri_tiles instead
  const uint32_t *tri_tiles;    // packed tile range per triangle (VB200_TILES_DEAD = dead)
  uint32_t list_cap;
  uint32_t num_tris;
  // pending ClearTarget()s folded into this launch: bit 0 colour, bit 1 depth. The kernel then takes the
  // attachment's prior contents from these constants instead of loading them and writes EVERY pixel of
  // every tile (including tiles no triangle touches), so no separate clear kernel runs.
  uint32_t clear_flags, clear_color;
  float clear_depth;
  // sort-first exchange fused into the write-back: the colour image of every OTHER rank (peer-mapped
  // device addresses over NVLink). Each pixel this rank produces is stored locally and to all peers,
  // so when every rank's tile kernel has finished, every rank holds the complete image.
  uint32_t num_peers;
  uint32_t *peer_color[7];
  // ... or ONE store to a multicast (NVLS) mapping of the image: the NVSwitch replicates it to every rank
  uint32_t *mc_color;
  uint32_t *color;
  float *depth;
  const float4 *interps;
  const float *unorm;    // 256 floats: float(i) / 255.0f, the exact quotients (texture_sampling.cpp:121-133, rasterizer.cpp:595-599)
  Vb200DrawCounters *counters;
  Vb200RasterState rs;
};

namespace vb200
{
// every launcher returns the number of kernel launches it issued (for vb200_stats::kernel_launches)
int launch_clear_u32(uint32_t *dst, uint32_t value, size_t count, cudaStream_t s);
int launch_clear_u8(uint8_t *dst, uint8_t value, size_t count, cudaStream_t s);
int launch_index_range(const void *ib, uint32_t index_type, uint32_t first, uint32_t count, uint32_t *range,
                       cudaStream_t s);
int launch_setup(const Vb200SetupParams &p, cudaStream_t s);
int launch_sort(uint32_t *list, const uint32_t *tile_count, uint32_t list_cap, uint32_t rank, uint32_t world,
                uint32_t ntiles, cudaStream_t s);
int launch_sample(const Vb200Image &img, int cube, uint64_t byte_offset, const float *uvw, float4 *out,
                  size_t count, cudaStream_t s);
int launch_tiles_pack(const uint32_t *color, uint32_t width, uint32_t height, uint32_t rank, uint32_t world,
                      uint32_t *dst, cudaStream_t s);
int launch_tiles_unpack(uint32_t *color, uint32_t width, uint32_t height, uint32_t world, uint32_t slots_per_rank,
                        const uint32_t *src, cudaStream_t s);
// sort-first multi-GPU plumbing (mgpu.cpp). Every rank holds the same symmetric buffers; peer[r] is rank r's copy
// mapped into this GPU's address space (peer[own rank] = the local copy).
struct Vb200PeerSet
{
  void *peer[8];
};
// cross-rank barrier on the stream: every rank writes `epoch` into slot `rank` of every rank's flag array, then
// waits until its own array shows `epoch` in every slot. Gives up after ~2 s and raises *timed_out (a peer that
// never arrives must not hang the GPU).
int launch_mgpu_barrier(const Vb200PeerSet &flags, uint32_t rank, uint32_t world, uint32_t epoch, uint32_t *timed_out,
                        cudaStream_t s);
// copy bytes [offset, offset + bytes) of the local copy of a symmetric buffer to the same range on every other rank:
// one multicast store per 16 bytes when `multicast` (an NVSwitch mapping of the buffer) is given, else one store per peer
int launch_mgpu_push(const Vb200PeerSet &bufs, void *multicast, uint32_t rank, uint32_t world, uint64_t offset,
                     uint64_t bytes, cudaStream_t s);
}    // namespace vb200
