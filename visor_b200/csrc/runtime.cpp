// runtime.cpp — implementation of the C-ABI in include/visor_b200.h.
//
// Host side of the B200 draw path: owns the CUDA stream, the HBM mirrors of the application's host
// memory (VkDeviceMemory semantics, memory.cpp:5-41), the per-pipeline JIT cache
// (SPIR-V -> PTX -> ptxas -> cudaLibrary) and the launch sequence that replaces DrawTriangles
// (rasterizer.cpp:363-520):
//
//   [index range] -> K1 vertex (JIT) -> K2 setup + binning -> [per-tile sort] -> K4 tiles (JIT)
//
// There is no CPU fallback anywhere in this file: without a CUDA device every compute entry point
// returns VB200_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvPTXCompiler.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/visor_b200.h"
#include "device_types.h"
#include "kernels.h"
#include "runtime_internal.h"
#include "spirv_ptx.h"
#include "ptx_inline.h"

extern "C" const char vb200_scaffold_ptx[];
extern "C" const unsigned long long vb200_scaffold_ptx_size;

struct vb200_entry
{
  vb200::ShaderEntry e;
  vb200_shader *owner;
  uint64_t serial;
};
struct vb200_shader
{
  std::unique_ptr<vb200::ShaderModule> mod;
  std::vector<std::unique_ptr<vb200_entry>> entries;
};

namespace
{
std::string g_error;
uint64_t g_serial = 1;

int setError(int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

// The kernels of scaffold.cu. Each is compiled per shader function: K_VERTEX with a vertex entry, the
// tile kernels with a fragment entry, on first use.
enum
{
  K_VERTEX = 0,
  K_TILE_ORDERED = 1,
  K_TILE_RESOLVE = 2,    // + resolve mode 0..4
  K_COUNT = 7
};
const char *const kKernelNames[K_COUNT] = {
    "vb200_k_vertex",
    "vb200_k_tile_ordered",
    "vb200_k_tile_resolve_min_first",
    "vb200_k_tile_resolve_min_last",
    "vb200_k_tile_resolve_max_first",
    "vb200_k_tile_resolve_max_last",
    "vb200_k_tile_resolve_last_wins",
};
struct JitKernel
{
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  unsigned threads = 0;    // CTA size the kernel was written for (.maxntid of its entry)
};

// Set of disjoint half-open byte intervals [lo, hi) of a mirror, kept sorted and merged.
struct IntervalSet
{
  std::vector<std::pair<size_t, size_t>> v;    // (lo, hi)
  bool empty() const { return v.empty(); }
  void clear() { v.clear(); }
  void add(size_t lo, size_t hi)
  {
    if(lo >= hi)
      return;
    std::vector<std::pair<size_t, size_t>> out;
    out.reserve(v.size() + 1);
    size_t i = 0;
    for(; i < v.size() && v[i].second < lo; i++)
      out.push_back(v[i]);
    for(; i < v.size() && v[i].first <= hi; i++)    // overlapping or touching: absorb
    {
      lo = std::min(lo, v[i].first);
      hi = std::max(hi, v[i].second);
    }
    out.push_back({lo, hi});
    for(; i < v.size(); i++)
      out.push_back(v[i]);
    v.swap(out);
  }
  bool covers(size_t lo, size_t hi) const    // [lo, hi) lies inside one interval (intervals are merged)
  {
    for(auto &r : v)
      if(r.first <= lo && hi <= r.second)
        return true;
    return lo >= hi;
  }
  bool overlaps(size_t lo, size_t hi) const
  {
    for(auto &r : v)
      if(r.first < hi && lo < r.second)
        return true;
    return false;
  }
  // removes [lo, hi); returns whether anything was removed
  bool remove(size_t lo, size_t hi)
  {
    bool hit = false;
    std::vector<std::pair<size_t, size_t>> out;
    for(auto &r : v)
    {
      if(r.second <= lo || r.first >= hi)
      {
        out.push_back(r);
        continue;
      }
      hit = true;
      if(r.first < lo)
        out.push_back({r.first, lo});
      if(r.second > hi)
        out.push_back({hi, r.second});
    }
    v.swap(out);
    return hit;
  }
  // the parts of [lo, hi) that lie in neither this set nor `other`, in order
  std::vector<std::pair<size_t, size_t>> gaps(size_t lo, size_t hi, const IntervalSet &other) const
  {
    IntervalSet u = *this;
    for(auto &r : other.v)
      u.add(r.first, r.second);
    std::vector<std::pair<size_t, size_t>> out;
    size_t at = lo;
    for(auto &r : u.v)
    {
      if(r.second <= at)
        continue;
      if(r.first >= hi)
        break;
      if(r.first > at)
        out.push_back({at, r.first});
      at = std::max(at, r.second);
    }
    if(at < hi)
      out.push_back({at, hi});
    return out;
  }
};

struct Mirror
{
  uint8_t *host = nullptr;
  size_t size = 0;
  uint8_t *dev = nullptr;
  bool pinned = false;       // page-locked by vb200_mem_register (caller guarantees lifetime)
  bool explicitReg = false;
  bool deviceLocal = false;    // DEVICE_LOCAL memory: the mirror is authoritative, no per-epoch upload/download
  uint64_t lastUse = 0;
  // coherent mode, this epoch (= since the last flush): bytes the device copy already holds from the host
  // (uploaded) and bytes device work produced, which a flush brings back (written). A read uploads exactly
  // the bytes of its range that are in neither set, so results produced earlier in the same submit (a
  // rendered or copied-to image that is sampled afterwards, a partially copied buffer) are never overwritten
  // with stale host data.
  IntervalSet uploaded, written;
};

template <typename T>
struct DevBuf
{
  T *p = nullptr;
  size_t cap = 0;
  bool reserve(size_t n)
  {
    if(n <= cap)
      return true;
    if(p)
      cudaFree(p);
    p = nullptr;
    size_t want = std::max(n, cap + cap / 2);
    if(cudaMalloc((void **)&p, want * sizeof(T)) != cudaSuccess)
    {
      cap = 0;
      return false;
    }
    cap = want;
    return true;
  }
  void release()
  {
    if(p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// A batch of recorded draws (see kernels.h: Vb200BatchDraw). The key is everything that must be uniform across
// one launch of the vertex, setup and tile kernels; it is compared bytewise.
struct BatchKey
{
  Vb200Env env;
  uint64_t vs, fs;    // serials of the shader entries
  uint32_t topology, frontFace, cullMode, depthOp, depthWrite, blend[4];
  uint8_t *colorDev, *depthDev;
  uint32_t width, height, vertexBound, nslots;
  int tileKernelId;
};
struct Batch
{
  enum SpanKind
  {
    SPAN_HOST,       // vertices [spanBase, spanBase + spanCount): non-indexed draws, ranges read back
    SPAN_INDEXED,    // the same, measured from host-readable indices; may still be widened to SPAN_ALL
    SPAN_ALL,        // indexed, references every bound vertex several times: all bound vertices
    SPAN_DEVICE,     // indexed, range only measurable on the device
  };
  struct Draw
  {
    Vb200BatchDraw dev;
    int kind;
    uint32_t spanBase, spanCount, usedVerts;
    const uint8_t *ibHost;    // SPAN_INDEXED: the draw's indices in host memory
  };
  bool active = false, closed = false, ran = false;
  BatchKey key;
  cudaKernel_t kVertex = nullptr, kTile = nullptr;
  unsigned tileThreads = 0;
  uint32_t clearFlags = 0, clearColorWord = 0;
  float clearDepthValue = 0.0f;
  std::vector<Draw> draws;
  uint64_t numTris = 0, numVerts = 0;
  void reset()
  {
    active = closed = ran = false;
    kVertex = kTile = nullptr;
    tileThreads = 0;
    clearFlags = clearColorWord = 0;
    clearDepthValue = 0.0f;
    draws.clear();
    numTris = numVerts = 0;
  }
};

struct Context
{
  Batch batch;
  bool ready = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  int syncMode = VB200_SYNC_COHERENT;
  uint64_t epoch = 1;
  uint64_t mirrorMerges = 0;    // createMirror calls that replaced existing mirrors
  std::map<uintptr_t, Mirror> mirrors;
  std::vector<uint8_t *> zombies;    // device memory of merged mirrors, freed at the next flush
  std::map<std::pair<uint64_t, int>, JitKernel> kernels;    // (shader entry serial, K_*) -> loaded kernel
  // scratch
  DevBuf<Vb200RasterVertex> rv;
  DevBuf<float4> interps;
  DevBuf<Vb200TriRecord> setup;
  DevBuf<uint32_t> tileCount, list, triTiles;
  uint32_t *range = nullptr;                // device {min,max}
  DevBuf<Vb200BatchDraw> batchDraws;        // device tables of a batch of several draws
  DevBuf<Vb200VertexSpan> batchSpans;
  int64_t optBatchDraws = 1;                // 0: every draw is rasterised by itself, as the reference replays them
  // ClearTarget()s not yet executed: device address of the attachment -> fill word and pixel count.
  // The next draw into the attachment folds them into its tile kernel; anything else that touches the
  // memory (another reader, a download) materialises them with the fill kernel first.
  struct PendingClear
  {
    uint32_t value;
    size_t count;
  };
  std::map<uint8_t *, PendingClear> pendingClears;
  int64_t optFuseClears = 1;
  int64_t optSlotKeys = 1;    // 0: never put the record slot into the visibility key (the path of draws >= 2^24 triangles)
  int64_t optTileListCap = 0; // > 0: entries per tile list (testing aid: a tiny value forces the tile kernels' fallback scan)
  // fused sort-first exchange: a device range on this rank that holds colour targets -> the same range on the
  // peers (mapped here) and/or an NVSwitch multicast mapping of it
  struct ExchangeRange
  {
    size_t bytes = 0;
    bool exact = false;    // caller-managed association (vb200_set_peer_targets): only a target that starts here
    std::vector<uint8_t *> peers;
    uint8_t *multicast = nullptr;
  };
  std::map<uint8_t *, ExchangeRange> exchange;
  int64_t optMgpuMirrors = 0;    // mirrors of registered host ranges are symmetric buffers (collective registration)
  Vb200DrawCounters *counters = nullptr;    // device
  float *unorm = nullptr;                   // device: float(i) / 255.0f for every byte value
  const char *lastTileKernel = "";
  // present path: copies on a second stream, ordered against the library stream with events
  cudaStream_t copyStream = nullptr;
  struct Present
  {
    cudaEvent_t ready = nullptr, done = nullptr;
    const uint8_t *dev = nullptr;
    size_t size = 0;
    bool active = false;
  } presents[8];
  int activePresents = 0;
  vb200_stats stats;
  uint32_t ownerRank = 0, ownerWorld = 1;
  int64_t optRasterPath = 0, optCountFragments = 0, optTimeKernels = 0;
  int stickyCuda = 0;
  cudaEvent_t userEvents[16] = {};
  cudaEvent_t phaseEvents[8] = {};    // 0..4 draw stages, 5..6 clears
  double phaseMs[VB200_PHASES] = {};
  uint64_t phaseCount[VB200_PHASES] = {};
} g;

// phase timing (diagnostic, enabled by option "time_kernels"): CUDA events on the library stream
// around each stage of a draw; accumulated per phase.
int phaseMark(int i)
{
  if(!g.optTimeKernels)
    return VB200_OK;
  if(!g.phaseEvents[i] && cudaEventCreate(&g.phaseEvents[i]) != cudaSuccess)
    return setError(VB200_ERR_CUDA, "cudaEventCreate failed");
  if(cudaEventRecord(g.phaseEvents[i], g.stream) != cudaSuccess)
    return setError(VB200_ERR_CUDA, "cudaEventRecord failed");
  return VB200_OK;
}
void phaseAccumulate(int phase, int a, int b)
{
  if(!g.optTimeKernels || !g.phaseEvents[a] || !g.phaseEvents[b])
    return;
  float ms = 0.0f;
  cudaEventSynchronize(g.phaseEvents[b]);
  if(cudaEventElapsedTime(&ms, g.phaseEvents[a], g.phaseEvents[b]) == cudaSuccess)
  {
    g.phaseMs[phase] += ms;
    g.phaseCount[phase]++;
  }
  else
    cudaGetLastError();
}

#define CU(call)                                                                                       \
  do                                                                                                   \
  {                                                                                                    \
    cudaError_t _e = (call);                                                                           \
    if(_e != cudaSuccess)                                                                              \
    {                                                                                                  \
      g.stickyCuda = 1;                                                                                \
      return setError(VB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                                       \
    }                                                                                                  \
  } while(0)

int flushBatch();    // launches the kernels of the draws recorded so far (below, next to vb200_draw)
int flushBatchKeepOpen();

int requireReadyKeepBatch()
{
  if(!g.ready)
  {
    int rc = vb200_init(0);
    if(rc != VB200_OK)
      return rc;
  }
  if(g.stickyCuda)
    return setError(VB200_ERR_CUDA, "a previous CUDA error is sticky: %s", g_error.c_str());
  return VB200_OK;
}

// Every entry point but vb200_draw starts here: whatever it does is ordered after the draws recorded before it.
int requireReady()
{
  int rc = requireReadyKeepBatch();
  return rc ? rc : flushBatch();
}

void freeMirrorMemory(uint8_t *dev)
{
  if(!dev)
    return;
  if(vb200::mgpu_is_symmetric(dev))
    vb200::mgpu_sym_free(dev);
  else
    cudaFree(dev);
}

bool isDevicePointer(const void *p)
{
  cudaPointerAttributes a;
  if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int materializeClears(const uint8_t *dev, size_t bytes);

// ---- residency -------------------------------------------------------------------------------
Mirror *findMirror(const void *host, size_t size)
{
  uintptr_t a = (uintptr_t)host;
  auto it = g.mirrors.upper_bound(a);
  if(it == g.mirrors.begin())
    return nullptr;
  --it;
  Mirror &m = it->second;
  if(a >= (uintptr_t)m.host && a + size <= (uintptr_t)m.host + m.size)
    return &m;
  return nullptr;
}

int createMirror(void *host, size_t size, bool pin, Mirror **out)
{
  // merge with anything the new range overlaps (aliased buffers of one allocation)
  uintptr_t lo = (uintptr_t)host, hi = lo + size;
  std::vector<uintptr_t> victims;
  for(auto &kv : g.mirrors)
  {
    uintptr_t mlo = (uintptr_t)kv.second.host, mhi = mlo + kv.second.size;
    if(mlo < hi && lo < mhi)
    {
      victims.push_back(kv.first);
      lo = std::min(lo, mlo);
      hi = std::max(hi, mhi);
    }
  }
  // Draws recorded into the open batch but not launched yet hold device addresses inside the mirrors about
  // to be replaced (attachments included): they run first, so that the copies below carry their results over.
  if(!victims.empty())
  {
    if(int frc = flushBatchKeepOpen())
      return frc;
    g.mirrorMerges++;    // vb200_draw resolves its ranges again when this moved under it
  }
  Mirror nm;
  nm.host = (uint8_t *)lo;
  nm.size = hi - lo;
  const size_t devBytes = std::max<size_t>(nm.size + 16, 256);
  if(pin && g.optMgpuMirrors && vb200::mgpu_active())
  {
    // sort-first mode: the mirror of a registered range is a symmetric buffer (same allocation on every rank,
    // all of them mapped here), so that colour targets inside it get the fused exchange and inputs can be
    // uploaded in slices. Collective: every rank registers the same ranges in the same order.
    if(int arc = vb200::mgpu_sym_alloc(devBytes, &nm.dev))
      return arc;
  }
  else
    CU(cudaMalloc((void **)&nm.dev, devBytes));
  // 16 bytes of slack behind every mirror, zeroed: texel fetches of formats narrower than 4 bytes read 4 bytes
  // per texel and may look up to 3 bytes past the end of an image that ends the mirror
  CU(cudaMemsetAsync(nm.dev + nm.size, 0, devBytes - nm.size, g.stream));
  for(uintptr_t key : victims)
  {
    Mirror &old = g.mirrors[key];
    if(int mrc = materializeClears(old.dev, old.size))    // deferred clears are keyed by the old address
      return mrc;
    // keep device-side contents (attachments may hold results not yet downloaded)
    CU(cudaMemcpyAsync(nm.dev + ((uintptr_t)old.host - lo), old.dev, old.size, cudaMemcpyDeviceToDevice, g.stream));
    const size_t shift = (uintptr_t)old.host - lo;
    for(auto &w : old.written.v)
      nm.written.add(w.first + shift, w.second + shift);
    for(auto &u : old.uploaded.v)
      nm.uploaded.add(u.first + shift, u.second + shift);
    if(old.pinned)
      cudaHostUnregister(old.host);
    // Launches already enqueued (and the draw being assembled right now) may hold device addresses
    // inside the old mirror: it stays allocated until the next flush has drained the stream.
    g.zombies.push_back(old.dev);
    nm.explicitReg |= old.explicitReg;
    nm.deviceLocal |= old.explicitReg && old.deviceLocal;
    g.mirrors.erase(key);
  }
  if(pin)
  {
    if(cudaHostRegister(nm.host, nm.size, cudaHostRegisterDefault) == cudaSuccess)
      nm.pinned = true;
    else
      cudaGetLastError();    // pageable copies still work
    nm.explicitReg = true;
  }
  nm.lastUse = g.epoch;
  auto ins = g.mirrors.emplace(lo, nm);
  *out = &ins.first->second;
  return VB200_OK;
}

int phaseMark(int i);
void phaseAccumulate(int phase, int a, int b);

// run the deferred clears that overlap [dev, dev+bytes)
int materializeClears(const uint8_t *dev, size_t bytes)
{
  for(auto it = g.pendingClears.begin(); it != g.pendingClears.end();)
  {
    const uint8_t *lo = it->first, *hi = lo + it->second.count * 4;
    if(lo < dev + bytes && dev < hi)
    {
      phaseMark(5);
      g.stats.kernel_launches += vb200::launch_clear_u32((uint32_t *)it->first, it->second.value, it->second.count, g.stream);
      phaseMark(6);
      phaseAccumulate(VB200_PHASE_CLEAR, 5, 6);
      it = g.pendingClears.erase(it);
    }
    else
      ++it;
  }
  CU(cudaGetLastError());
  return VB200_OK;
}

enum Access
{
  ACC_READ = 1,           // kernel reads host-authored data: upload on first use per epoch (coherent mode)
  ACC_WRITE = 2,          // kernel writes: download at flush (coherent mode)
  ACC_OVERWRITE = 4,      // the whole range is overwritten first: no upload needed
  ACC_KEEP_PENDING = 8,   // caller deals with deferred clears of the range itself
};

// host (or device) pointer -> device pointer usable by kernels
int resolveInner(const void *ptr, size_t size, int access, uint8_t **out);

// host pointer (or device pointer) -> device address, with the upload / write bookkeeping of the sync
// mode; work that will write the range is ordered after present copies still reading it
int resolve(const void *ptr, size_t size, int access, uint8_t **out)
{
  int rc = resolveInner(ptr, size, access, out);
  if(rc == VB200_OK && g.activePresents && (access & (ACC_WRITE | ACC_OVERWRITE)))
    for(auto &pr : g.presents)
      if(pr.active && pr.dev < *out + size && *out < pr.dev + pr.size)
        CU(cudaStreamWaitEvent(g.stream, pr.done, 0));
  return rc;
}

int resolveInner(const void *ptr, size_t size, int access, uint8_t **out)
{
  *out = nullptr;
  if(!ptr)
    return setError(VB200_ERR_INVALID, "NULL resource pointer");
  if(isDevicePointer(ptr))
  {
    *out = (uint8_t *)ptr;
    if(!(access & ACC_KEEP_PENDING) && !g.pendingClears.empty())
      return materializeClears(*out, size);
    return VB200_OK;
  }
  Mirror *m = findMirror(ptr, size);
  if(!m)
  {
    int rc = createMirror((void *)ptr, size, false, &m);
    if(rc)
      return rc;
  }
  m->lastUse = g.epoch;
  const size_t off = (uintptr_t)ptr - (uintptr_t)m->host;
  if(g.syncMode == VB200_SYNC_COHERENT && !m->deviceLocal)
  {
    // (a frame of many draws resolves the same ranges over and over: the common case is "already there")
    if((access & ACC_READ) && !(access & ACC_OVERWRITE) && !m->uploaded.covers(off, off + size) &&
       !m->written.covers(off, off + size))
    {
      for(auto &gap : m->uploaded.gaps(off, off + size, m->written))
      {
        CU(cudaMemcpyAsync(m->dev + gap.first, m->host + gap.first, gap.second - gap.first, cudaMemcpyHostToDevice,
                           g.stream));
        g.stats.h2d_bytes += gap.second - gap.first;
        m->uploaded.add(gap.first, gap.second);
      }
    }
    if((access & (ACC_WRITE | ACC_OVERWRITE)) && !m->written.covers(off, off + size))
      m->written.add(off, off + size);
  }
  *out = m->dev + off;
  if(!(access & ACC_KEEP_PENDING) && !g.pendingClears.empty())
    return materializeClears(*out, size);
  return VB200_OK;
}

// ---- JIT -------------------------------------------------------------------------------------
// ---- kernel + shader -> cubin --------------------------------------------------------------------
// scaffold.cu is shipped as PTX. A shader function (vb200_vs / vb200_fs, PTX from spirv_ptx.cpp) is
// spliced into the text of the ONE kernel that calls it and the module is compiled whole by ptxas
// (the nvPTXCompiler library, linked statically: no toolkit or driver JIT is needed at run time). Compared with linking a prebuilt cubin against the shader, the call disappears: ptxas inlines
// the shader, schedules its loads across the kernel's unrolled pixel rows, turns its reads of the
// kernel-parameter environment into constant-bank loads and allocates registers for the whole (the
// C3 resolve kernel drops from 54 to 45 registers, the vertex kernel from 62 + a stack frame to 34).
struct ScaffoldText
{
  bool parsed = false, ok = false;
  std::string prologue;    // .version/.target, shared-memory declarations
  std::string helpers;     // vb200_fetch_attr, vb200_sample_tex, vb200_sample_cube
  std::string entries[K_COUNT];
  std::string epilogue;    // .file table + .section .debug_str of the -lineinfo .loc directives
} scaffold;

// removes every `.extern .func ... ;` prototype (the definitions are spliced in instead)
void stripExternFuncs(std::string &t)
{
  for(size_t at = t.find(".extern .func"); at != std::string::npos; at = t.find(".extern .func", at))
  {
    const size_t end = t.find(';', at);
    if(end == std::string::npos)
      break;
    t.erase(at, end + 1 - at);
  }
}

bool parseScaffold()
{
  if(scaffold.parsed)
    return scaffold.ok;
  scaffold.parsed = true;
  std::string t(vb200_scaffold_ptx, (size_t)vb200_scaffold_ptx_size);
  if(const char *path = getenv("VB200_SCAFFOLD_PTX"))
  {
    // tuning aid: kernel scaffolds from a PTX file (a variant of scaffold.cu built with other -D switches,
    // `make -C visor_b200 variants`) instead of the embedded text
    if(FILE *f = fopen(path, "rb"))
    {
      t.clear();
      char buf[65536];
      size_t got;
      while((got = fread(buf, 1, sizeof(buf), f)) > 0)
        t.append(buf, got);
      fclose(f);
      fprintf(stderr, "visor_b200: kernel scaffolds from %s (%zu bytes)\n", path, t.size());
    }
  }
  while(!t.empty() && t.back() == '\0')
    t.pop_back();
  stripExternFuncs(t);
  // trailer of the module: the .file table and .debug_str section the -lineinfo .loc directives use
  size_t dbg = t.find("\n\t.file\t");
  if(dbg == std::string::npos)
    dbg = t.find("\n\t.section\t.debug_str");
  if(dbg != std::string::npos)
  {
    scaffold.epilogue = t.substr(dbg + 1);
    t.erase(dbg + 1);
  }
  const size_t firstEntry = t.find("\n.visible .entry ");
  size_t firstFunc = std::string::npos;
  for(const char *pat : {"\n.visible .func", "\n.func", "\n.weak .func"})
    firstFunc = std::min(firstFunc, t.find(pat));
  if(firstEntry == std::string::npos || firstFunc == std::string::npos || firstFunc > firstEntry)
    return false;
  scaffold.prologue = t.substr(0, firstFunc + 1);
  scaffold.helpers = t.substr(firstFunc + 1, firstEntry - firstFunc);
  for(size_t at = firstEntry; at != std::string::npos;)
  {
    const size_t next = t.find("\n.visible .entry ", at + 1);
    const size_t nameAt = at + strlen("\n.visible .entry ");
    const std::string name = t.substr(nameAt, t.find('(', nameAt) - nameAt);
    const std::string body = t.substr(at + 1, (next == std::string::npos ? t.size() : next) - at);
    if(body.find("\n.func") != std::string::npos || body.find("\n.visible .func") != std::string::npos)
      return false;    // layout assumption broken: a function defined between kernels
    for(int k = 0; k < K_COUNT; k++)
      if(name == kKernelNames[k])
        scaffold.entries[k] = body;
    at = next;
  }
  for(int k = 0; k < K_COUNT; k++)
    if(scaffold.entries[k].empty())
      return false;
  scaffold.ok = true;
  return true;
}

// One ptxas run over a complete module. `spills` receives the spill-store bytes ptxas reports.
int runPtxas(const std::string &text, const char *name, std::vector<char> &cubin, unsigned *spills)
{
  nvPTXCompilerHandle h;
  if(nvPTXCompilerCreate(&h, text.size(), text.c_str()) != NVPTXCOMPILE_SUCCESS)
    return setError(VB200_ERR_LINK, "nvPTXCompilerCreate failed");
  std::string maxreg;
  const char *opts[5] = {"--gpu-name=sm_100a", "--generate-line-info", "--verbose", nullptr, nullptr};
  int nopts = 3;
  if(const char *mr = getenv("VB200_JIT_MAXRREGCOUNT"))    // tuning aid
  {
    maxreg = std::string("--maxrregcount=") + mr;
    opts[nopts++] = maxreg.c_str();
  }
  const nvPTXCompileResult r = nvPTXCompilerCompile(h, nopts, opts);
  if(r != NVPTXCOMPILE_SUCCESS)
  {
    std::string log;
    size_t n = 0;
    if(nvPTXCompilerGetErrorLogSize(h, &n) == NVPTXCOMPILE_SUCCESS && n > 0)
    {
      log.resize(n + 1);
      nvPTXCompilerGetErrorLog(h, &log[0]);
    }
    nvPTXCompilerDestroy(&h);
    return setError(VB200_ERR_LINK, "ptxas failed on %s (%d): %s", name, (int)r, log.c_str());
  }
  if(spills)
  {
    // "... N bytes stack frame, N bytes spill stores, N bytes spill loads"
    *spills = 0;
    std::string log;
    size_t n = 0;
    if(nvPTXCompilerGetInfoLogSize(h, &n) == NVPTXCOMPILE_SUCCESS && n > 0)
    {
      log.resize(n + 1);
      nvPTXCompilerGetInfoLog(h, &log[0]);
      for(size_t at = log.find(" bytes spill stores"); at != std::string::npos; at = log.find(" bytes spill stores", at + 1))
      {
        size_t d = at;
        while(d > 0 && isdigit((unsigned char)log[d - 1]))
          d--;
        *spills += (unsigned)strtoul(log.c_str() + d, nullptr, 10);
      }
    }
  }
  size_t sz = 0;
  nvPTXCompilerGetCompiledProgramSize(h, &sz);
  cubin.resize(sz);
  nvPTXCompilerGetCompiledProgram(h, cubin.data());
  nvPTXCompilerDestroy(&h);
  return VB200_OK;
}

// CTA size a kernel of the scaffold was written for: the .maxntid directive of its entry (__launch_bounds__)
unsigned kernelThreads(int which)
{
  if(!parseScaffold())
    return 0;
  const std::string &entry = scaffold.entries[which];
  const size_t at = entry.find(".maxntid ");
  return at == std::string::npos ? 0u : (unsigned)strtoul(entry.c_str() + at + strlen(".maxntid "), nullptr, 10);
}

// ptxas step: kernel text + shader function -> one sm_100a cubin. Needs no device.
int compileKernel(int which, const vb200_entry *shader, std::vector<char> &cubin)
{
  if(!parseScaffold())
    return setError(VB200_ERR_LINK, "embedded kernel scaffold PTX has an unexpected layout");
  std::string body = shader->e.ptx;
  for(const char *directive : {".version", ".target", ".address_size"})
  {
    const size_t at = body.find(directive);
    if(at != std::string::npos)
      body.erase(at, body.find('\n', at) - at);
  }
  stripExternFuncs(body);
  {
    // line info for the profiler: the shader's instructions carry no .loc of their own, so after inlining they
    // would be booked on whatever kernel line happens to precede them in the cubin. They are pinned to line 1 of
    // scaffold.cu instead (tools/ncu_regions.py, tools/sass_lines.py list that line as "shader, inlined").
    const size_t f = scaffold.epilogue.find("scaffold.cu\"");
    const size_t dir = f == std::string::npos ? f : scaffold.epilogue.rfind(".file", f);
    const size_t open = body.find("\n{\n");
    if(dir != std::string::npos && open != std::string::npos)
    {
      const unsigned long idx = strtoul(scaffold.epilogue.c_str() + dir + 5, nullptr, 10);
      size_t at = open + 3;
      while(body.compare(at, 6, "  .reg") == 0)    // after the register declarations
        at = body.find('\n', at) + 1;
      if(idx)
        body.insert(at, "  .loc " + std::to_string(idx) + " 1 0\n");
    }
  }
  // Tile kernels: a bound on the resident CTAs per SM (.minnctapersm) makes ptxas fit the register
  // allocation granule instead of landing just above it. Whether a shader leaves room for that is only
  // known after inlining, so the bound is dropped when ptxas reports more than a few spilled words.
  // VB200_JIT_MINCTAS overrides (tuning aid).
  const unsigned kSpillTolerance = 64;
  const unsigned threads = kernelThreads(which);
  // resident CTAs asked of ptxas: 40 warps per SM for the ordered kernel (48 registers), 32 for the resolve
  // kernels (64 registers: measured on B200, C3 tile kernel 138.9 -> 135.5 us and C5 878 -> 832 us against
  // 40 warps at 48 registers with a few spilled words; 48 warps at 40 registers: 144 / 975 us)
  int minCtas = which == K_VERTEX || !threads ? 0 : (int)((which == K_TILE_ORDERED ? 1280u : 1024u) / threads);
  bool forced = false;
  if(const char *mc = getenv("VB200_JIT_MINCTAS"))
  {
    minCtas = which != K_VERTEX ? atoi(mc) : 0;
    forced = true;
  }
  // ptxas keeps `.func` calls as calls. The shader entry point, and the helpers it calls (attribute fetch, texture
  // unit), are therefore inlined into the kernel on the PTX text (ptx_inline.cpp), which is what lets ptxas
  // schedule the shader's loads across the kernel's own arithmetic and allocate registers for the whole.
  // VB200_JIT_NO_INLINE=1 keeps the calls (A/B measurements).
  std::string inlinedEntry = scaffold.entries[which];
  bool needBody = true, needHelpers = true;
  if(!getenv("VB200_JIT_NO_INLINE"))
  {
    int serial = 0;
    const char *fn = which == K_VERTEX ? "vb200_vs" : "vb200_fs";
    vb200::ptx_inline_calls(inlinedEntry, body, fn, &serial);
    needHelpers = false;
    for(const char *helper : {"vb200_fetch_attr", "vb200_sample_cube", "vb200_sample_tex"})
    {
      // (the texture unit is ~2 800 lines of PTX per copy: a shader with many sample instructions keeps the call,
      // or the module would take ptxas seconds)
      size_t sites = 0;
      for(size_t at = inlinedEntry.find(helper); at != std::string::npos; at = inlinedEntry.find(helper, at + 1))
        sites++;
      if(sites <= 12)
        vb200::ptx_inline_calls(inlinedEntry, scaffold.helpers, helper, &serial);
      needHelpers |= vb200::ptx_has_call(inlinedEntry, helper);
    }
    needBody = vb200::ptx_has_call(inlinedEntry, fn);
    needHelpers |= needBody;
  }
  for(;;)
  {
    std::string text;
    text.reserve(scaffold.prologue.size() + scaffold.helpers.size() + body.size() + inlinedEntry.size() +
                 scaffold.epilogue.size() + 64);
    text += scaffold.prologue;
    if(needHelpers)
      text += scaffold.helpers;
    if(needBody)
      text += body;
    text += "\n";
    std::string entry = inlinedEntry;
    const size_t at = entry.find(".maxntid ");
    const size_t eol = at == std::string::npos ? at : entry.find('\n', at);
    if(minCtas > 0 && eol != std::string::npos)
      entry.insert(eol, "\n.minnctapersm " + std::to_string(minCtas));
    text += entry;
    text += scaffold.epilogue;
    unsigned spills = 0;
    if(const char *prefix = getenv("VB200_DUMP_PTX"))
    {
      // developer aid: the module as ptxas gets it: <prefix>.<kernel>.ptx
      const std::string path = std::string(prefix) + "." + kKernelNames[which] + ".ptx";
      if(FILE *f = fopen(path.c_str(), "wb"))
      {
        fwrite(text.data(), 1, text.size(), f);
        fclose(f);
      }
    }
    int rc = runPtxas(text, kKernelNames[which], cubin, &spills);
    if(rc)
      return rc;
    if(getenv("VB200_JIT_VERBOSE"))
      fprintf(stderr, "visor_b200 jit: %s minctas=%d spill stores=%u B cubin=%zu B\n", kKernelNames[which], minCtas,
              spills, cubin.size());
    if(spills <= kSpillTolerance || minCtas == 0 || forced)
      break;
    minCtas = 0;    // the occupancy bound costs real spills with this shader: let ptxas choose
  }
  if(const char *prefix = getenv("VB200_DUMP_CUBIN"))
  {
    // developer aid: keep the cubin for cuobjdump -sass / -res-usage: <prefix>.<kernel>.cubin
    const std::string path = std::string(prefix) + "." + kKernelNames[which] + ".cubin";
    if(FILE *f = fopen(path.c_str(), "wb"))
    {
      fwrite(cubin.data(), 1, cubin.size(), f);
      fclose(f);
    }
  }
  return VB200_OK;
}

// the loaded kernel `which` specialised for `shader`, compiled on first use
int getKernel(int which, const vb200_entry *shader, cudaKernel_t *out, unsigned *threads = nullptr)
{
  const auto key = std::make_pair(shader->serial, which);
  auto it = g.kernels.find(key);
  if(it != g.kernels.end())
  {
    *out = it->second.kernel;
    if(threads)
      *threads = it->second.threads;
    return VB200_OK;
  }
  std::vector<char> cubin;
  int rc = compileKernel(which, shader, cubin);
  if(rc)
    return rc;
  JitKernel jk;
  cudaError_t e = cudaLibraryLoadData(&jk.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if(e != cudaSuccess)
    return setError(VB200_ERR_LINK, "cudaLibraryLoadData failed: %s", cudaGetErrorString(e));
  e = cudaLibraryGetKernel(&jk.kernel, jk.lib, kKernelNames[which]);
  if(e != cudaSuccess)
  {
    cudaLibraryUnload(jk.lib);
    return setError(VB200_ERR_LINK, "cudaLibraryGetKernel(%s) failed: %s", kKernelNames[which], cudaGetErrorString(e));
  }
  jk.threads = kernelThreads(which);
  g.kernels.emplace(key, jk);
  *out = jk.kernel;
  if(threads)
    *threads = jk.threads;
  return VB200_OK;
}

// interpolant slots a (VS, FS) pair moves per vertex
uint32_t pipelineSlots(const vb200_entry *vs, const vb200_entry *fs)
{
  const uint32_t mask = vs->e.out_slot_mask | fs->e.in_slot_mask;
  uint32_t nslots = 1;
  for(uint32_t s = 0; s < VB200_MAX_SLOTS; s++)
    if(mask & (1u << s))
      nslots = s + 1;
  return nslots;
}

uint32_t formatBytes(uint32_t fmt)
{
  switch(fmt)
  {
    case 109: case 107: case 108: return 16;
    case 106: case 104: case 105: return 12;
    case 103: case 101: case 102: return 8;
    case 100: case 98: case 99: return 4;
    case 37: return 4;
    default: return 0;
  }
}

uint64_t sliceBytes(const vb200_image &im)
{
  // CalcSubresourceByteOffset (precompiled.cpp:18-33): full mip chain of one layer
  uint32_t mw = im.width, mh = im.height;
  uint64_t slice = 0;
  for(uint32_t m = 0; m < im.mip_levels; m++)
  {
    slice += (uint64_t)(mw * mh * im.bytes_per_pixel);
    mw = std::max(1u, mw >> 1);
    mh = std::max(1u, mh >> 1);
  }
  return slice;
}

uint64_t imageBytes(const vb200_image &im)
{
  // vkGetImageMemoryRequirements (images.cpp:51-65)
  uint64_t sz = (uint64_t)im.width * im.height * std::max(1u, im.array_layers) * im.bytes_per_pixel;
  if(im.mip_levels > 1)
    sz *= 2;
  return sz;
}

// Bytes of a sampled image to make resident. Texel fetches of formats narrower than 4 bytes read 4 bytes
// per texel (texture_sampling.cpp:121-133), so the last texels look up to 3 bytes past the image, at
// whatever follows it in the same allocation. Those bytes are mirrored only when they are known to exist:
// when the image lies inside an already mirrored (registered) range that extends that far. The library
// never reads host memory beyond what the caller described.
size_t sampledExtent(const void *pixels, uint64_t bytes)
{
  return (size_t)bytes + (findMirror(pixels, (size_t)bytes + 4) ? 4 : 0);
}

const vb200_binding *findBinding(const vb200_draw_state *s, uint32_t set, uint32_t binding)
{
  for(uint32_t i = 0; i < s->num_bindings; i++)
    if(s->bindings[i].set == set && s->bindings[i].binding == binding)
      return &s->bindings[i];
  return nullptr;
}

int fillResources(const vb200_draw_state *s, const vb200::ShaderEntry &e, Vb200Env &env)
{
  for(const vb200::ResourceSlot &r : e.resources)
  {
    const vb200_binding *b = findBinding(s, r.set, r.binding);
    if(!b)
      return setError(VB200_ERR_INVALID, "shader reads descriptor (set %u, binding %u) which is not bound", r.set,
                      r.binding);
    if(r.is_image)
    {
      if(!b->is_image)
        return setError(VB200_ERR_INVALID, "descriptor (%u,%u) is not an image", r.set, r.binding);
      const vb200_image &im = b->image;
      if((im.width & 3) || (im.height & 3))
        return setError(VB200_ERR_INVALID, "texture size must be a multiple of 4 (texture_sampling.cpp:123-133)");
      uint8_t *dev;
      int rc = resolve(im.pixels, sampledExtent(im.pixels, imageBytes(im)), ACC_READ, &dev);
      if(rc)
        return rc;
      Vb200Image &d = env.images[r.slot];
      d.pixels = dev;
      d.width = im.width;
      d.height = im.height;
      d.bpp = im.bytes_per_pixel;
      d.format = im.format;
      d.mips = im.mip_levels;
      d.layers = im.array_layers;
      d.slice_bytes = sliceBytes(im);
    }
    else
    {
      if(b->is_image)
        return setError(VB200_ERR_INVALID, "descriptor (%u,%u) is not a buffer", r.set, r.binding);
      if(b->offset > b->buffer.size)
        return setError(VB200_ERR_INVALID, "descriptor offset beyond buffer");
      uint8_t *dev;
      int rc = resolve(b->buffer.bytes, b->buffer.size, ACC_READ, &dev);
      if(rc)
        return rc;
      if(((uintptr_t)(dev + b->offset)) & 15)
        return setError(VB200_ERR_INVALID, "uniform buffer address must be 16-byte aligned");
      env.res[r.slot] = dev + b->offset;
    }
  }
  return VB200_OK;
}

// `dependent`: the kernel begins with griddepcontrol.wait and only consumes what the kernel launched right before
// it produced (setup after vertex, tiles after setup/sort): it is launched with programmatic stream
// serialization, so that its CTAs are scheduled while the predecessor's last wave drains instead of after a
// launch gap (the predecessor executes griddepcontrol.launch_dependents when it starts). VB200_NO_PDL=1: plain
// stream order (A/B measurements).
int launchKernel(cudaKernel_t k, dim3 grid, dim3 block, void **args, bool dependent = false)
{
  static bool pdl = getenv("VB200_NO_PDL") == nullptr;
  if(dependent && pdl)
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = g.stream;
    cudaLaunchAttribute attr = {};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)k, args);
    if(e == cudaErrorNotSupported || e == cudaErrorInvalidValue)
    {
      // an environment without programmatic launches (some virtualised or instrumented set-ups): plain order
      cudaGetLastError();
      pdl = false;
      CU(cudaLaunchKernel((const void *)k, grid, block, args, 0, g.stream));
    }
    else
      CU(e);
  }
  else
    CU(cudaLaunchKernel((const void *)k, grid, block, args, 0, g.stream));
  g.stats.kernel_launches++;
  return VB200_OK;
}

int checkTarget(const vb200_image *im, const char *what)
{
  if(!im || !im->pixels)
    return setError(VB200_ERR_INVALID, "%s: no image memory", what);
  if(im->width == 0 || im->height == 0 || im->width > 8192 || im->height > 8192)
    return setError(VB200_ERR_INVALID, "%s: extent %ux%u outside 1..8192", what, im->width, im->height);
  return VB200_OK;
}
}    // namespace

namespace vb200
{
int set_error(int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}
cudaStream_t library_stream()
{
  if(g.ready)
    flushBatch();    // whoever enqueues on the stream next comes after the recorded draws
  return g.ready ? g.stream : nullptr;
}
int library_device()
{
  return g.device;
}
void count_launches(int n)
{
  g.stats.kernel_launches += (uint64_t)n;
}
void set_exchange_range(uint8_t *local, size_t bytes, const std::vector<uint8_t *> &peers, uint8_t *multicast)
{
  Context::ExchangeRange &r = g.exchange[local];
  r.bytes = bytes;
  r.exact = false;
  r.peers = peers;
  r.multicast = multicast;
}
void clear_exchange_range(uint8_t *local)
{
  g.exchange.erase(local);
}
}    // namespace vb200

// ---- batches of draws -------------------------------------------------------------------------
// (declared here, next to their only writer; Context holds one)
namespace
{
// Launches the kernels for draws [first, last) of the recorded batch as ONE vertex / setup / tile pass.
int runBatch(size_t first, size_t last, bool foldClears)
{
  Batch &b = g.batch;
  const BatchKey &k = b.key;
  int rc;
  // ---- vertex spans. Indexed draws that together reference at least as many indices as the bound buffers
  // hold vertices (meshes; many sub-mesh draws out of one buffer) share ONE span over all bound vertices: no
  // pass over the index buffer, every vertex shaded once for the whole batch. Otherwise each draw keeps the
  // span measured for it.
  uint64_t idxVerts = 0, idxSpan = 0;
  bool wantAll = false, deviceRange = false;
  for(size_t i = first; i < last; i++)
  {
    const Batch::Draw &d = b.draws[i];
    if(d.kind == Batch::SPAN_ALL)
      wantAll = true;
    if(d.kind != Batch::SPAN_HOST)
      idxVerts += d.usedVerts;
    if(d.kind == Batch::SPAN_DEVICE)
      deviceRange = true;
  }
  if(!wantAll && idxVerts < k.vertexBound)
    for(size_t i = first; i < last; i++)
    {
      Batch::Draw &d = b.draws[i];
      if(d.kind != Batch::SPAN_INDEXED)
        continue;
      uint32_t lo = 0xffffffffu, hi = 0;
      if(d.dev.index_type == 0u)
        for(uint32_t j = 0; j < d.usedVerts; j++)
        {
          uint16_t v;
          memcpy(&v, d.ibHost + 2 * (size_t)j, 2);
          lo = std::min<uint32_t>(lo, v);
          hi = std::max<uint32_t>(hi, v);
        }
      else
        for(uint32_t j = 0; j < d.usedVerts; j++)
        {
          uint32_t v;
          memcpy(&v, d.ibHost + 4 * (size_t)j, 4);
          lo = std::min(lo, v);
          hi = std::max(hi, v);
        }
      d.spanBase = lo;
      // (indices at or beyond vertexBound kill their triangle in the setup kernel; nothing beyond is shaded)
      d.spanCount = lo < k.vertexBound ? std::min(hi, k.vertexBound - 1u) - lo + 1u : 0u;
      idxSpan += d.spanCount;
    }
  if(deviceRange && last - first > 1)
  {
    // index data only the device can read: worth one shared span when the draws reference enough of the buffer
    if(wantAll || idxVerts >= k.vertexBound / 4u)
      wantAll = true;
  }
  else if(deviceRange)
    wantAll = wantAll || idxVerts >= k.vertexBound;
  else if(idxVerts >= k.vertexBound || idxSpan >= k.vertexBound)
    wantAll = true;
  if(deviceRange && !wantAll && last - first > 1)
  {
    // several draws whose ranges only the device can measure and that use a small part of a large buffer:
    // one at a time (each measures its own range on the device)
    for(size_t i = first; i < last; i++)
      if((rc = runBatch(i, i + 1, foldClears && i == first)))
        return rc;
    return VB200_OK;
  }
  std::vector<Vb200VertexSpan> spans;
  std::vector<Vb200BatchDraw> draws;
  spans.reserve(last - first);
  draws.reserve(last - first);
  uint32_t records = 0, numTris = 0;
  int allSpan = -1;
  const uint32_t *rangeDev = nullptr;
  for(size_t i = first; i < last; i++)
  {
    const Batch::Draw &d = b.draws[i];
    Vb200BatchDraw dd = d.dev;
    dd.tri_base = numTris;
    numTris += dd.num_tris;
    const bool all = d.kind == Batch::SPAN_ALL || (wantAll && d.kind != Batch::SPAN_HOST);
    if(all)
    {
      if(allSpan < 0)
      {
        allSpan = (int)spans.size();
        spans.push_back({0u, k.vertexBound, records, 0u});
        records += k.vertexBound;
      }
      dd.span = (uint32_t)allSpan;
    }
    else if(d.kind == Batch::SPAN_DEVICE)
    {
      // alone in its batch: span0 = [min, max] of its indices (k_index_range), at most vertexBound records
      g.stats.kernel_launches += vb200::launch_index_range(dd.ib, dd.index_type, dd.first, d.usedVerts, g.range, g.stream);
      rangeDev = g.range;
      dd.span = (uint32_t)spans.size();
      spans.push_back({0u, k.vertexBound, records, 0u});
      records += k.vertexBound;
    }
    else
    {
      // a span equal to the previous one is shared (the same vertices drawn again)
      if(!spans.empty() && (int)spans.size() - 1 != allSpan && spans.back().src_base == d.spanBase &&
         spans.back().count == d.spanCount)
        dd.span = (uint32_t)spans.size() - 1u;
      else
      {
        dd.span = (uint32_t)spans.size();
        spans.push_back({d.spanBase, d.spanCount, records, 0u});
        records += d.spanCount;
      }
    }
    draws.push_back(dd);
  }
  if(records == 0 || numTris == 0)
    return VB200_OK;

  // ---- scratch
  const uint32_t W = k.width, H = k.height;
  const uint32_t tilesX = (W + VB200_TILE - 1) / VB200_TILE, tilesY = (H + VB200_TILE - 1) / VB200_TILE;
  const uint32_t ntiles = tilesX * tilesY;
  const uint32_t ownedTiles = (ntiles + g.ownerWorld - 1u) / g.ownerWorld;
  // Entries per tile list: a power of two of at least four times the average load (a perspective mesh puts
  // several times the average into its far tiles), within [256, 4096]. It is only a performance knob: a
  // tile that receives more triangles than that is rasterised from the packed tile ranges instead.
  uint32_t listCap = 256;
  {
    const uint64_t want = 4ull * numTris / std::max(1u, ntiles) + 64;
    while(listCap < want && listCap < 4096u)
      listCap <<= 1;
    if(g.optTileListCap > 0)
      listCap = (uint32_t)std::min<int64_t>(g.optTileListCap, 1 << 20);
  }
  const bool tables = draws.size() > 1;
  if(!g.rv.reserve(records) || !g.interps.reserve((size_t)records * k.nslots) || !g.setup.reserve(numTris) ||
     !g.triTiles.reserve(numTris) || !g.list.reserve((size_t)ownedTiles * listCap) || !g.tileCount.reserve(ntiles) ||
     (tables && (!g.batchDraws.reserve(draws.size()) || !g.batchSpans.reserve(spans.size()))))
  {
    g.stickyCuda = 1;
    return setError(VB200_ERR_CUDA, "out of device memory for draw scratch");
  }
  if(tables)
  {
    // (pageable source: the copy is staged by the driver before the call returns)
    CU(cudaMemcpyAsync(g.batchDraws.p, draws.data(), draws.size() * sizeof(Vb200BatchDraw), cudaMemcpyHostToDevice, g.stream));
    CU(cudaMemcpyAsync(g.batchSpans.p, spans.data(), spans.size() * sizeof(Vb200VertexSpan), cudaMemcpyHostToDevice, g.stream));
  }

  // ---- K1: vertex stage
  phaseMark(0);
  Vb200VertexParams vp;
  memset(&vp, 0, sizeof(vp));
  vp.spans = (tables && spans.size() > 1) ? g.batchSpans.p : nullptr;
  vp.num_spans = (uint32_t)spans.size();
  vp.span0 = spans[0];
  vp.range = rangeDev;
  vp.count = records;
  vp.vertex_bound = k.vertexBound;
  vp.rv = g.rv.p;
  vp.interps = g.interps.p;
  vp.nslots = k.nslots;
  vp.width = W;
  vp.height = H;
  vp.tile_count = g.tileCount.p;
  vp.tile_count_n = ntiles;
  Vb200Env env = k.env;
  {
    void *args[] = {&env, &vp};
    if((rc = launchKernel(b.kVertex, dim3((records + 127) / 128), dim3(128), args)))
      return rc;
  }

  // ---- K2: setup + binning
  phaseMark(1);
  Vb200SetupParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.draws = tables ? g.batchDraws.p : nullptr;
  sp.spans = tables ? g.batchSpans.p : nullptr;
  sp.num_draws = (uint32_t)draws.size();
  sp.draw0 = draws[0];
  sp.span0 = spans[0];
  sp.range = rangeDev;
  sp.num_tris = numTris;
  sp.topology = k.topology;
  sp.vertex_bound = k.vertexBound;
  sp.rv = g.rv.p;
  sp.tri = g.setup.p;
  sp.tri_tiles = g.triTiles.p;
  sp.tile_count = g.tileCount.p;
  sp.list = g.list.p;
  sp.list_cap = listCap;
  sp.counters = g.counters;
  sp.front_face = k.frontFace;
  sp.cull_mode = k.cullMode;
  sp.width = W;
  sp.height = H;
  sp.tiles_x = tilesX;
  sp.tiles_y = tilesY;
  sp.owner_rank = g.ownerRank;
  sp.owner_world = g.ownerWorld;
  g.stats.kernel_launches += vb200::launch_setup(sp, g.stream);    // (the vertex kernel zeroed the tile counters)
  phaseMark(2);

  g.lastTileKernel = kKernelNames[k.tileKernelId];
  Vb200TileParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.tri = g.setup.p;
  tp.rv = g.rv.p;
  tp.list = g.list.p;
  tp.list_cap = listCap;
  tp.tile_count = g.tileCount.p;
  tp.tri_tiles = g.triTiles.p;
  tp.num_tris = numTris;
  tp.color = (uint32_t *)k.colorDev;
  tp.depth = (float *)k.depthDev;
  tp.interps = g.interps.p;
  tp.counters = g.counters;
  tp.rs.width = W;
  tp.rs.height = H;
  tp.rs.tiles_x = tilesX;
  tp.rs.tiles_y = tilesY;
  tp.rs.tiles_x_magic = (uint32_t)(0x100000000ull / tilesX) + 1u;
  tp.unorm = g.unorm;
  tp.rs.depth_op = k.depthOp;
  tp.rs.depth_write = k.depthWrite;
  tp.rs.has_depth = k.depthDev ? 1u : 0u;
  tp.rs.blend_enable = k.blend[0];
  tp.rs.src_factor = k.blend[1];
  tp.rs.dst_factor = k.blend[2];
  tp.rs.blend_op = k.blend[3];
  tp.rs.nslots = k.nslots;
  tp.rs.owner_rank = g.ownerRank;
  tp.rs.owner_world = g.ownerWorld;
  tp.rs.count_fragments = g.optCountFragments ? 1u : 0u;
  tp.rs.color_bpp = 4;
  tp.rs.slot_keys = (g.optSlotKeys && numTris < (1u << 24) - 1u) ? 1u : 0u;
  if(foldClears)
  {
    tp.clear_flags = b.clearFlags;
    tp.clear_color = b.clearColorWord;
    tp.clear_depth = b.clearDepthValue;
  }
  if(g.ownerWorld > 1)
  {
    auto it = g.exchange.upper_bound(k.colorDev);
    if(it != g.exchange.begin())
    {
      --it;
      const size_t off = (size_t)(k.colorDev - it->first);
      if(it->second.exact ? off == 0 : off + (size_t)W * H * 4 <= it->second.bytes)
      {
        if(it->second.multicast)
          tp.mc_color = (uint32_t *)(it->second.multicast + off);
        else
        {
          tp.num_peers = (uint32_t)std::min<size_t>(it->second.peers.size(), 7);
          for(uint32_t r = 0; r < tp.num_peers; r++)
            tp.peer_color[r] = (uint32_t *)(it->second.peers[r] + off);
        }
      }
    }
  }

  // ---- K3' + K4. The ordered path needs each tile's list in submission order (the appends of the setup
  // kernel arrive in any order); nothing here waits for the device.
  if(k.tileKernelId == K_TILE_ORDERED)
    g.stats.kernel_launches += vb200::launch_sort(g.list.p, g.tileCount.p, listCap, g.ownerRank, g.ownerWorld, ntiles,
                                                  g.stream);
  phaseMark(3);
  {
    void *args[] = {&env, &tp};
    if((rc = launchKernel(b.kTile, dim3(ownedTiles), dim3(b.tileThreads), args, true)))
      return rc;
  }
  phaseMark(4);
  phaseAccumulate(VB200_PHASE_VERTEX, 0, 1);
  phaseAccumulate(VB200_PHASE_SETUP, 1, 2);
  phaseAccumulate(VB200_PHASE_BIN, 2, 3);
  phaseAccumulate(VB200_PHASE_TILES, 3, 4);
  CU(cudaGetLastError());
  return VB200_OK;
}

// runs what has been recorded; the batch stays open for more draws with the same key
int flushBatchKeepOpen()
{
  Batch &b = g.batch;
  if(!b.active || b.draws.empty())
    return VB200_OK;
  const int rc = runBatch(0, b.draws.size(), !b.ran);
  b.ran = true;
  b.draws.clear();
  b.numTris = b.numVerts = 0;
  return rc;
}

int flushBatch()
{
  Batch &b = g.batch;
  if(!b.active)
    return VB200_OK;
  b.active = false;    // (first: runBatch's helpers must not re-enter)
  int rc = VB200_OK;
  if(!b.draws.empty())
    rc = runBatch(0, b.draws.size(), !b.ran);
  else if(!b.ran && b.clearFlags)
  {
    // the clears this batch took over were never folded into a tile kernel: put them back
    if(b.clearFlags & 1u)
      g.pendingClears[b.key.colorDev] = {b.clearColorWord, (size_t)b.key.width * b.key.height};
    if(b.clearFlags & 2u)
    {
      uint32_t bits;
      memcpy(&bits, &b.clearDepthValue, 4);
      g.pendingClears[b.key.depthDev] = {bits, (size_t)b.key.width * b.key.height};
    }
  }
  b.reset();
  return rc;
}
}    // namespace

// =================================================================================================
extern "C" {

const char *vb200_last_tile_kernel(void)
{
  return g.lastTileKernel;
}

int vb200_abi_version(void)
{
  return VB200_ABI_VERSION;
}

const char *vb200_last_error(void)
{
  return g_error.c_str();
}

int vb200_init(int device)
{
  if(g.ready)
    return VB200_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess || n <= 0)
  {
    cudaGetLastError();
    return setError(VB200_ERR_NO_DEVICE, "no CUDA device available (%s); visor_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if(device < 0 || device >= n)
    return setError(VB200_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  CU(cudaSetDevice(device));
  g.device = device;
  CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&g.copyStream, cudaStreamNonBlocking));
  CU(cudaMalloc((void **)&g.range, 2 * sizeof(uint32_t)));
  {
    float table[256];
    for(int i = 0; i < 256; i++)
      table[i] = (float)i / 255.0f;    // IEEE single division, as the reference's float(byte) / 255.0f
    CU(cudaMalloc((void **)&g.unorm, sizeof(table)));
    CU(cudaMemcpy(g.unorm, table, sizeof(table), cudaMemcpyHostToDevice));
  }
  CU(cudaMalloc((void **)&g.counters, sizeof(Vb200DrawCounters)));
  CU(cudaMemsetAsync(g.counters, 0, sizeof(Vb200DrawCounters), g.stream));
  memset(&g.stats, 0, sizeof(g.stats));
  g.ready = true;
  g.stickyCuda = 0;
  // tuning aid: VB200_OPTIONS="name=value,name=value" applies vb200_set_option() settings at start-up
  if(const char *opts = getenv("VB200_OPTIONS"))
  {
    std::string all(opts);
    size_t at = 0;
    while(at < all.size())
    {
      size_t end = all.find(',', at);
      if(end == std::string::npos)
        end = all.size();
      const std::string item = all.substr(at, end - at);
      const size_t eq = item.find('=');
      if(eq != std::string::npos)
      {
        int rc = vb200_set_option(item.substr(0, eq).c_str(), strtoll(item.c_str() + eq + 1, nullptr, 0));
        if(rc)
          return rc;
      }
      at = end + 1;
    }
  }
  return VB200_OK;
}

void vb200_shutdown(void)
{
  if(!g.ready)
    return;
  flushBatch();
  cudaStreamSynchronize(g.stream);
  for(auto &kv : g.kernels)
    cudaLibraryUnload(kv.second.lib);
  g.kernels.clear();
  for(auto &kv : g.mirrors)
  {
    if(kv.second.pinned)
      cudaHostUnregister(kv.second.host);
    freeMirrorMemory(kv.second.dev);
  }
  g.mirrors.clear();
  g.rv.release();
  g.interps.release();
  g.setup.release();
  g.tileCount.release();
  g.list.release();
  cudaFree(g.range);
  g.triTiles.release();
  g.batchDraws.release();
  g.batchSpans.release();
  cudaFree(g.counters);
  cudaFree(g.unorm);
  for(auto &pr : g.presents)
  {
    if(pr.ready)
      cudaEventDestroy(pr.ready);
    if(pr.done)
      cudaEventDestroy(pr.done);
    pr.ready = pr.done = nullptr;
    pr.active = false;
  }
  g.activePresents = 0;
  cudaStreamDestroy(g.copyStream);
  cudaStreamDestroy(g.stream);
  g.ready = false;
}

void *vb200_stream(void)
{
  if(g.ready)
    flushBatch();    // work the caller enqueues comes after the recorded draws
  return g.ready ? (void *)g.stream : nullptr;
}

// ---- shaders ---------------------------------------------------------------------------------
vb200_shader *vb200_shader_create(const uint32_t *code, size_t words)
{
  if(!code || words < 5)
  {
    setError(VB200_ERR_SPIRV, "empty SPIR-V module");
    return nullptr;
  }
  std::string err;
  vb200::ShaderModule *m = vb200::compile_spirv(code, words, &err);
  if(!m)
  {
    setError(VB200_ERR_SPIRV, "CompileFunction: %s", err.c_str());
    return nullptr;
  }
  vb200_shader *s = new vb200_shader;
  s->mod.reset(m);
  for(vb200::ShaderEntry &e : m->entries)
  {
    std::unique_ptr<vb200_entry> en(new vb200_entry);
    en->e = e;
    en->owner = s;
    en->serial = g_serial++;
    s->entries.push_back(std::move(en));
  }
  return s;
}

vb200_entry *vb200_shader_entry(vb200_shader *shader, const char *name)
{
  if(!shader || !name)
    return nullptr;
  for(auto &e : shader->entries)
    if(e->e.name == name)
      return e.get();
  setError(VB200_ERR_INVALID, "GetFuncPointer: no entry point named '%s'", name);
  return nullptr;
}

void vb200_shader_destroy(vb200_shader *shader)
{
  if(!shader)
    return;
  if(g.ready)
    flushBatch();    // recorded draws may still need them
  // drop the kernels compiled for this module's entries
  for(auto &e : shader->entries)
    for(auto it = g.kernels.begin(); it != g.kernels.end();)
    {
      if(it->first.first == e->serial)
      {
        if(g.ready)
        {
          cudaStreamSynchronize(g.stream);
          cudaLibraryUnload(it->second.lib);
        }
        it = g.kernels.erase(it);
      }
      else
        ++it;
    }
  delete shader;
}

int vb200_link_check(const vb200_entry *vs, const vb200_entry *fs, uint64_t *cubin_size)
{
  if(!vs || !fs || vs->e.stage != vb200::STAGE_VERTEX || fs->e.stage != vb200::STAGE_FRAGMENT)
    return setError(VB200_ERR_INVALID, "link_check needs one vertex and one fragment entry");
  // every kernel a draw with this pair can launch (a draw itself compiles only the ones it uses)
  uint64_t total = 0;
  for(int k = 0; k < K_COUNT; k++)
  {
    std::vector<char> cubin;
    int rc = compileKernel(k, k == K_VERTEX ? vs : fs, cubin);
    if(rc)
      return rc;
    total += cubin.size();
  }
  if(cubin_size)
    *cubin_size = total;
  return VB200_OK;
}

const char *vb200_entry_ptx(const vb200_entry *entry)
{
  return entry ? entry->e.ptx.c_str() : nullptr;
}

int vb200_entry_stage(const vb200_entry *entry)
{
  return entry ? entry->e.stage : -1;
}

int vb200_entry_num_resources(const vb200_entry *entry)
{
  return entry ? (int)entry->e.resources.size() : 0;
}

int vb200_entry_resource(const vb200_entry *entry, int index, uint32_t *set, uint32_t *binding, uint32_t *is_image)
{
  if(!entry || index < 0 || index >= (int)entry->e.resources.size())
    return setError(VB200_ERR_INVALID, "entry_resource: index out of range");
  const vb200::ResourceSlot &r = entry->e.resources[index];
  if(set)
    *set = r.set;
  if(binding)
    *binding = r.binding;
  if(is_image)
    *is_image = r.is_image ? 1u : 0u;
  return VB200_OK;
}

// ---- residency -------------------------------------------------------------------------------
int vb200_set_sync_mode(int mode)
{
  if(g.ready)
    flushBatch();
  if(mode != VB200_SYNC_COHERENT && mode != VB200_SYNC_EXPLICIT)
    return setError(VB200_ERR_INVALID, "unknown sync mode %d", mode);
  g.syncMode = mode;
  return VB200_OK;
}

int vb200_mem_register(void *host, uint64_t size)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(!host || !size)
    return setError(VB200_ERR_INVALID, "mem_register: empty range");
  if(isDevicePointer(host))
    return VB200_OK;
  Mirror *m = findMirror(host, size);
  if(m && m->explicitReg)
    return VB200_OK;    // a sub-range of an allocation that is already registered
  // A registration states the exact extent of a live allocation. Automatically created mirrors that
  // overlap it are caches of whatever used to live at these addresses (the library never learns that an
  // unregistered array was freed): they must not be promoted — their range, page-locking and residency
  // flags would then cover unrelated memory. Bring back what they still owe the host, then drop them.
  const uintptr_t lo = (uintptr_t)host, hi = lo + size;
  bool owed = false;
  for(auto &kv : g.mirrors)
  {
    const uintptr_t mlo = (uintptr_t)kv.second.host, mhi = mlo + kv.second.size;
    if(mlo < hi && lo < mhi && !kv.second.explicitReg && !kv.second.written.empty())
      owed = true;
  }
  if(owed && (rc = vb200_flush()))
    return rc;
  for(auto it = g.mirrors.begin(); it != g.mirrors.end();)
  {
    const uintptr_t mlo = (uintptr_t)it->second.host, mhi = mlo + it->second.size;
    if(mlo < hi && lo < mhi && !it->second.explicitReg)
    {
      if((rc = materializeClears(it->second.dev, it->second.size)))
        return rc;
      g.zombies.push_back(it->second.dev);    // launches in flight may still address it; freed at the next flush
      it = g.mirrors.erase(it);
    }
    else
      ++it;
  }
  return createMirror(host, size, true, &m);
}

int vb200_mem_unregister(void *host)
{
  if(g.ready)
    flushBatch();
  if(!g.ready)
    return VB200_OK;
  auto it = g.mirrors.find((uintptr_t)host);
  if(it == g.mirrors.end())
  {
    Mirror *m = findMirror(host, 1);
    if(!m)
      return VB200_OK;
    it = g.mirrors.find((uintptr_t)m->host);
  }
  cudaStreamSynchronize(g.stream);
  if(g.activePresents)
    cudaStreamSynchronize(g.copyStream);    // a present copy may still be reading this mirror
  for(auto pc = g.pendingClears.begin(); pc != g.pendingClears.end();)    // nobody will ever see them
    pc = (pc->first >= it->second.dev && pc->first < it->second.dev + it->second.size) ? g.pendingClears.erase(pc)
                                                                                         : std::next(pc);
  if(it->second.pinned)
    cudaHostUnregister(it->second.host);
  freeMirrorMemory(it->second.dev);
  g.mirrors.erase(it);
  return VB200_OK;
}

int vb200_mem_upload(const void *host, uint64_t size)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(isDevicePointer(host))
    return VB200_OK;
  Mirror *m = findMirror(host, size);
  if(!m)
  {
    rc = createMirror((void *)host, size, false, &m);
    if(rc)
      return rc;
  }
  const size_t off = (uintptr_t)host - (uintptr_t)m->host;
  CU(cudaMemcpyAsync(m->dev + off, host, size, cudaMemcpyHostToDevice, g.stream));
  g.stats.h2d_bytes += size;
  return VB200_OK;
}

int vb200_mgpu_upload(const void *host, uint64_t size)
{
  int rc = requireReady();
  if(rc)
    return rc;
  int rank = 0, world = 1;
  vb200_mgpu_info(&rank, &world, nullptr);
  Mirror *m = findMirror(host, size);
  if(world <= 1 || !m || !vb200::mgpu_is_symmetric(m->dev))
    return setError(VB200_ERR_INVALID, "mgpu_upload: the range is not mirrored in a symmetric buffer "
                                       "(vb200_mgpu_init, option mgpu_mirrors, vb200_mem_register)");
  const size_t off = (uintptr_t)host - (uintptr_t)m->host;
  if(off & 15)
    return setError(VB200_ERR_INVALID, "mgpu_upload: the range must start 16-byte aligned inside its registration");
  // every rank sends 1/world of the bytes over ITS PCIe link and replicates that slice to the other ranks over
  // NVLink; the few bytes left over by the division are uploaded by everybody
  const size_t slice = (size / (size_t)world) & ~(size_t)15;
  const uint8_t *src = (const uint8_t *)host;
  if(slice)
  {
    CU(cudaMemcpyAsync(m->dev + off + rank * slice, src + rank * slice, slice, cudaMemcpyHostToDevice, g.stream));
    g.stats.h2d_bytes += slice;
    if((rc = vb200_mgpu_push(m->dev + off + rank * slice, slice)))
      return rc;
  }
  const size_t done = slice * (size_t)world;
  if(done < size)
  {
    CU(cudaMemcpyAsync(m->dev + off + done, src + done, size - done, cudaMemcpyHostToDevice, g.stream));
    g.stats.h2d_bytes += size - done;
  }
  return VB200_OK;
}

int vb200_mem_download(void *host, uint64_t size)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(isDevicePointer(host))
    return VB200_OK;
  Mirror *m = findMirror(host, size);
  if(!m)
    return setError(VB200_ERR_INVALID, "mem_download: range is not mirrored");
  const size_t off = (uintptr_t)host - (uintptr_t)m->host;
  if((rc = materializeClears(m->dev + off, size)))
    return rc;
  CU(cudaMemcpyAsync(host, m->dev + off, size, cudaMemcpyDeviceToHost, g.stream));
  g.stats.d2h_bytes += size;
  return VB200_OK;
}

int vb200_mem_host_write(const void *host, uint64_t size)
{
  // The host is about to modify [host, host+size) in the middle of a submit (vkCmdCopyBuffer /
  // vkCmdCopyBufferToImage replayed as memcpy, cmd_exec.cpp:143-182): wait for in-flight uploads that
  // read the old contents, then forget that the range was uploaded so the next draw re-reads it.
  if(!g.ready || !host || !size || isDevicePointer(host))
    return VB200_OK;
  const uintptr_t lo = (uintptr_t)host, hi = lo + size;
  bool pending = false;
  for(auto &kv : g.mirrors)
  {
    Mirror &m = kv.second;
    const uintptr_t mlo = (uintptr_t)m.host, mhi = mlo + m.size;
    if(mlo >= hi || lo >= mhi)
      continue;
    const size_t rlo = (size_t)(std::max(lo, mlo) - mlo), rhi = (size_t)(std::min(hi, mhi) - mlo);
    pending |= m.uploaded.remove(rlo, rhi);
  }
  if(pending)
    CU(cudaStreamSynchronize(g.stream));
  return VB200_OK;
}

int vb200_mem_set_device_local(void *host, int device_local)
{
  int rc = requireReady();
  if(rc)
    return rc;
  Mirror *m = host ? findMirror(host, 1) : nullptr;
  if(!m || !m->explicitReg)
    return setError(VB200_ERR_INVALID, "mem_set_device_local: %p is not inside a registered range", host);
  m->deviceLocal = device_local != 0;
  return VB200_OK;
}

namespace
{
int copyBytes(const uint8_t *srcHost, uint8_t *dstHost, uint64_t size, const char *what)
{
  if(size == 0)
    return VB200_OK;
  if(!srcHost || !dstHost)
    return setError(VB200_ERR_INVALID, "%s: NULL buffer memory", what);
  uint8_t *srcDev, *dstDev;
  int rc = resolve(srcHost, size, ACC_READ, &srcDev);
  if(rc)
    return rc;
  // the whole destination range is overwritten: no upload of its old contents; marked written, so a
  // coherent-mode flush brings the copy back to host memory
  if((rc = resolve(dstHost, size, ACC_WRITE | ACC_OVERWRITE, &dstDev)))
    return rc;
  CU(cudaMemcpyAsync(dstDev, srcDev, size, cudaMemcpyDeviceToDevice, g.stream));
  return VB200_OK;
}
}    // namespace

int vb200_copy_buffer(const vb200_buffer *src, uint64_t src_offset, const vb200_buffer *dst, uint64_t dst_offset,
                      uint64_t size)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(!src || !dst)
    return setError(VB200_ERR_INVALID, "copy_buffer: NULL buffer");
  if(src_offset + size > src->size || dst_offset + size > dst->size)
    return setError(VB200_ERR_INVALID, "copy_buffer: range outside the buffer");
  return copyBytes((const uint8_t *)src->bytes + src_offset, (uint8_t *)dst->bytes + dst_offset, size, "copy_buffer");
}

int vb200_copy_buffer_to_image(const vb200_buffer *src, uint64_t buffer_offset, const vb200_image *dst,
                               uint32_t mip_level, uint32_t array_layer)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(!src || !dst || !dst->pixels)
    return setError(VB200_ERR_INVALID, "copy_buffer_to_image: NULL argument");
  if(mip_level >= std::max(1u, dst->mip_levels) || array_layer >= std::max(1u, dst->array_layers))
    return setError(VB200_ERR_INVALID, "copy_buffer_to_image: subresource (%u, %u) outside the image", mip_level,
                    array_layer);
  // CalcSubresourceByteOffset (precompiled.cpp:3-36): mips of the layer before this one + whole layers
  uint32_t mw = dst->width, mh = dst->height;
  uint64_t offs = 0;
  for(uint32_t m = 0; m < mip_level; m++)
  {
    offs += (uint64_t)(mw * mh * dst->bytes_per_pixel);
    mw = std::max(1u, mw >> 1);
    mh = std::max(1u, mh >> 1);
  }
  offs += sliceBytes(*dst) * array_layer;
  const uint32_t w = std::max(1u, dst->width >> mip_level), h = std::max(1u, dst->height >> mip_level);
  const uint64_t bytes = (uint64_t)w * h * dst->bytes_per_pixel;
  if(buffer_offset + bytes > src->size)
    return setError(VB200_ERR_INVALID, "copy_buffer_to_image: source range outside the buffer");
  return copyBytes((const uint8_t *)src->bytes + buffer_offset, (uint8_t *)dst->pixels + offs, bytes,
                   "copy_buffer_to_image");
}

int vb200_present(const vb200_image *image, void *dst_host, uint64_t dst_size, int *ticket)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if((rc = checkTarget(image, "present")))
    return rc;
  if(!dst_host || !ticket)
    return setError(VB200_ERR_INVALID, "present: NULL destination or ticket");
  const size_t bytes = (size_t)image->width * image->height * image->bytes_per_pixel;
  if(dst_size < bytes)
    return setError(VB200_ERR_INVALID, "present: destination holds %llu bytes, the image needs %zu",
                    (unsigned long long)dst_size, bytes);
  int slot = -1;
  for(int i = 0; i < 8 && slot < 0; i++)
    if(!g.presents[i].active)
      slot = i;
  if(slot < 0)    // every slot in flight: retire the oldest
  {
    slot = 0;
    CU(cudaEventSynchronize(g.presents[0].done));
    g.presents[0].active = false;
    g.activePresents--;
  }
  auto &pr = g.presents[slot];
  if(!pr.ready)
  {
    CU(cudaEventCreateWithFlags(&pr.ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&pr.done, cudaEventDisableTiming));
  }
  uint8_t *dev;
  if((rc = resolve(image->pixels, bytes, ACC_READ, &dev)))    // also materialises a clear nobody drew over
    return rc;
  CU(cudaEventRecord(pr.ready, g.stream));
  CU(cudaStreamWaitEvent(g.copyStream, pr.ready, 0));
  CU(cudaMemcpyAsync(dst_host, dev, bytes, cudaMemcpyDeviceToHost, g.copyStream));
  CU(cudaEventRecord(pr.done, g.copyStream));
  pr.dev = dev;
  pr.size = bytes;
  pr.active = true;
  g.activePresents++;
  g.stats.d2h_bytes += bytes;
  *ticket = slot + 1;
  return VB200_OK;
}

int vb200_present_wait(int ticket)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(ticket < 1 || ticket > 8)
    return setError(VB200_ERR_INVALID, "present_wait: bad ticket %d", ticket);
  auto &pr = g.presents[ticket - 1];
  if(!pr.active)
    return VB200_OK;    // already retired (by a flush or by slot reuse)
  CU(cudaEventSynchronize(pr.done));
  pr.active = false;
  g.activePresents--;
  return VB200_OK;
}

void *vb200_mem_device_ptr(const void *host)
{
  if(!g.ready || !host)
    return nullptr;
  if(isDevicePointer(host))
    return (void *)host;
  Mirror *m = findMirror(host, 1);
  return m ? (void *)(m->dev + ((uintptr_t)host - (uintptr_t)m->host)) : nullptr;
}

int vb200_flush(void)
{
  int rc = requireReady();
  if(rc)
    return rc;
  while(!g.pendingClears.empty())
  {
    auto it = g.pendingClears.begin();
    if((rc = materializeClears(it->first, it->second.count * 4)))
      return rc;
  }
  if(g.optMgpuMirrors && vb200::mgpu_active())
  {
    // sort-first with symmetric mirrors (the ICD's mode): every rank's tile kernels have stored their pixels
    // into all ranks' colour targets; one cross-rank barrier and every rank holds, and downloads, the whole image
    if((rc = vb200_mgpu_barrier()))
      return rc;
  }
  if(g.syncMode == VB200_SYNC_COHERENT)
  {
    for(auto &kv : g.mirrors)
    {
      Mirror &m = kv.second;
      for(auto &w : m.written.v)
      {
        CU(cudaMemcpyAsync(m.host + w.first, m.dev + w.first, w.second - w.first, cudaMemcpyDeviceToHost, g.stream));
        g.stats.d2h_bytes += w.second - w.first;
      }
    }
  }
  if(g.optMgpuMirrors && vb200::mgpu_active())
  {
    // ... and nobody starts storing the next frame into a peer that is still copying this one out
    if((rc = vb200_mgpu_barrier()))
      return rc;
  }
  CU(cudaStreamSynchronize(g.stream));
  if(g.activePresents)
  {
    CU(cudaStreamSynchronize(g.copyStream));
    for(auto &pr : g.presents)
      pr.active = false;
    g.activePresents = 0;
  }
  cudaError_t e = cudaGetLastError();
  if(e != cudaSuccess)
  {
    g.stickyCuda = 1;
    return setError(VB200_ERR_CUDA, "asynchronous CUDA error: %s", cudaGetErrorString(e));
  }
  for(uint8_t *z : g.zombies)
    freeMirrorMemory(z);
  g.zombies.clear();
  // new epoch: host memory is authoritative again
  size_t autoBytes = 0;
  for(auto &kv : g.mirrors)
  {
    kv.second.written.clear();
    kv.second.uploaded.clear();
    if(!kv.second.explicitReg)
      autoBytes += kv.second.size;
  }
  g.epoch++;
  if(autoBytes > ((size_t)8 << 30))
  {
    // drop auto-created mirrors that have not been touched for a while
    for(auto it = g.mirrors.begin(); it != g.mirrors.end();)
    {
      if(!it->second.explicitReg && it->second.lastUse + 4 < g.epoch)
      {
        freeMirrorMemory(it->second.dev);
        it = g.mirrors.erase(it);
      }
      else
        ++it;
    }
  }
  return VB200_OK;
}

// ---- operators -------------------------------------------------------------------------------
int vb200_clear_color(const vb200_image *target, const float rgba[4])
{
  int rc = requireReady();
  if(rc)
    return rc;
  if((rc = checkTarget(target, "ClearTarget")))
    return rc;
  if(!rgba)
    return setError(VB200_ERR_INVALID, "ClearTarget: NULL colour");
  // rasterizer.cpp:341-345: byte(f * 255.0f) per channel (x86 truncating convert, low 8 bits), B,G,R,A order
  auto toByte = [](float f) -> uint32_t {
    float v = f * 255.0f;
    int i = (v > -2147483648.0f && v < 2147483648.0f) ? (int)v : (int)0x80000000;
    return (uint32_t)i & 0xffu;
  };
  const uint32_t r = toByte(rgba[0]), gch = toByte(rgba[1]), b = toByte(rgba[2]), a = toByte(rgba[3]);
  const size_t px = (size_t)target->width * target->height;
  uint8_t *dev;
  if(target->bytes_per_pixel == 4)
  {
    if((rc = resolve(target->pixels, px * 4, ACC_OVERWRITE | ACC_KEEP_PENDING, &dev)))
      return rc;
    g.pendingClears[dev] = {b | (gch << 8) | (r << 16) | (a << 24), px};
    if(!g.optFuseClears && (rc = materializeClears(dev, px * 4)))
      return rc;
  }
  else if(target->bytes_per_pixel == 1)
  {
    if((rc = resolve(target->pixels, px, ACC_OVERWRITE, &dev)))
      return rc;
    g.stats.kernel_launches += vb200::launch_clear_u8(dev, (uint8_t)r, px, g.stream);
  }
  // other bpp: the reference does nothing (rasterizer.cpp:347-360)
  CU(cudaGetLastError());
  return VB200_OK;
}

int vb200_clear_depth(const vb200_image *target, float depth)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if((rc = checkTarget(target, "ClearTarget")))
    return rc;
  if(target->bytes_per_pixel != 4)
    return setError(VB200_ERR_INVALID, "depth clear requires 4 bytes per pixel (assert rasterizer.cpp:321)");
  const size_t px = (size_t)target->width * target->height;
  uint8_t *dev;
  if((rc = resolve(target->pixels, px * 4, ACC_OVERWRITE | ACC_KEEP_PENDING, &dev)))
    return rc;
  uint32_t bits;
  memcpy(&bits, &depth, 4);
  g.pendingClears[dev] = {bits, px};
  if(!g.optFuseClears && (rc = materializeClears(dev, px * 4)))
    return rc;
  return VB200_OK;
}

int vb200_draw(const vb200_draw_state *s, int num_verts, uint32_t first, int indexed)
{
  int rc = requireReadyKeepBatch();
  if(rc)
    return rc;
  if(!s || !s->pipeline)
    return setError(VB200_ERR_INVALID, "DrawTriangles: no pipeline bound");
  const vb200_pipeline *pl = s->pipeline;
  if(!pl->vs || !pl->fs)
    return setError(VB200_ERR_INVALID, "DrawTriangles: pipeline has no vertex/fragment shader");
  if(pl->vs->e.stage != vb200::STAGE_VERTEX || pl->fs->e.stage != vb200::STAGE_FRAGMENT)
    return setError(VB200_ERR_INVALID, "DrawTriangles: shader stages do not match");
  if((rc = checkTarget(&s->color, "DrawTriangles colour target")))
    return rc;
  if(s->color.bytes_per_pixel != 4)
    return setError(VB200_ERR_INVALID, "colour target must have 4 bytes per pixel");
  const bool hasDepth = s->depth.pixels != nullptr;
  if(hasDepth && (s->depth.width != s->color.width || s->depth.height != s->color.height ||
                  s->depth.bytes_per_pixel != 4))
    return setError(VB200_ERR_INVALID, "depth target must match the colour extent and be 32-bit");

  g.stats.draws++;
  // triangle count (rasterizer.cpp:133-136: whole triangles only; :152 strips need >= 3)
  uint32_t numTris = 0;
  if(pl->topology == 3u)
    numTris = num_verts >= 3 ? (uint32_t)num_verts / 3u : 0u;
  else if(pl->topology == 4u)
  {
    if(num_verts < 3)
      return setError(VB200_ERR_INVALID, "triangle strip needs >= 3 vertices (assert rasterizer.cpp:152)");
    numTris = (uint32_t)num_verts - 2u;
  }
  else
  {
    // "Unsupported primitive topology!" (rasterizer.cpp:235): draws nothing
    return VB200_OK;
  }
  g.stats.triangles_in += numTris;
  if(numTris == 0)
    return VB200_OK;
  const uint32_t usedVerts = pl->topology == 3u ? numTris * 3u : numTris + 2u;

  const uint32_t nslots = pipelineSlots(pl->vs, pl->fs);

  // ---- environment (device-side GPUState)
  Vb200Env env;
  memset(&env, 0, sizeof(env));
  uint32_t vertexBound = 0xffffffffu;    // how many vertices the bound buffers can hold
  for(uint32_t a = 0; a < 16; a++)
  {
    env.attrs[a].format = pl->vattrs[a].format;
    env.attrs[a].stride = pl->vattrs[a].stride;
    env.attrs[a].offset = pl->vattrs[a].offset;
    env.attrs[a].vb = pl->vattrs[a].vb;
    if(!(pl->vs->e.attr_mask & (1u << a)))
      continue;
    const uint32_t fb = formatBytes(pl->vattrs[a].format);
    if(!fb)
      return setError(VB200_ERR_INVALID, "Unhandled vertex attribute format %u at location %u (assert spirv_compile.cpp:625)",
                      pl->vattrs[a].format, a);
    if(pl->vattrs[a].vb >= 4)
      return setError(VB200_ERR_INVALID, "vertex binding %u out of range", pl->vattrs[a].vb);
    const auto &vb = s->vbs[pl->vattrs[a].vb];
    if(!vb.buffer.bytes)
      return setError(VB200_ERR_INVALID, "vertex buffer slot %u is not bound", pl->vattrs[a].vb);
    const uint64_t start = vb.offset + pl->vattrs[a].offset;
    if(start + fb > vb.buffer.size)
      return setError(VB200_ERR_INVALID, "vertex attribute %u lies outside its buffer", a);
    if(pl->vattrs[a].stride)
    {
      const uint64_t n = (vb.buffer.size - start - fb) / pl->vattrs[a].stride + 1;
      vertexBound = (uint32_t)std::min<uint64_t>(vertexBound, n);
    }
  }
  // A resolve below may replace mirrors by a merged one (a host range straddling older mirrors): device
  // addresses taken earlier in this draw would then point into the retired allocations, and what the draw
  // wrote to an attachment there would never reach the new mirror. Every range has a mirror after the
  // first pass, so a second pass cannot merge again.
  const uint64_t mergesBefore = g.mirrorMerges;
  bool resolvedTwice = false;
resolveRanges:
  for(uint32_t i = 0; i < 4; i++)
    if(s->vbs[i].buffer.bytes)
    {
      uint8_t *dev;
      if((rc = resolve(s->vbs[i].buffer.bytes, s->vbs[i].buffer.size, ACC_READ, &dev)))
        return rc;
      env.vb[i] = dev + s->vbs[i].offset;
    }
  if((rc = fillResources(s, pl->vs->e, env)))
    return rc;
  if((rc = fillResources(s, pl->fs->e, env)))
    return rc;
  memcpy(env.push, s->pushconsts, 128);

  // ---- index buffer
  const uint8_t *ibDev = nullptr;
  const uint8_t *ibHost = nullptr;    // the same indices in host memory, when the host copy is known to be current
  if(indexed)
  {
    if(!s->ib.buffer.bytes)
      return setError(VB200_ERR_INVALID, "indexed draw without an index buffer");
    const uint32_t isz = s->ib.index_type == 0u ? 2u : 4u;
    if(s->ib.offset + ((uint64_t)first + usedVerts) * isz > s->ib.buffer.size)
      return setError(VB200_ERR_INVALID, "index range lies outside the index buffer");
    uint8_t *dev;
    if((rc = resolve(s->ib.buffer.bytes, s->ib.buffer.size, ACC_READ, &dev)))
      return rc;
    ibDev = dev + s->ib.offset;
    if(((uintptr_t)ibDev) & (isz - 1))
      return setError(VB200_ERR_INVALID, "index buffer offset is not aligned to the index size");
    if(g.syncMode == VB200_SYNC_COHERENT && !isDevicePointer(s->ib.buffer.bytes))
    {
      // coherent mode: host memory is authoritative unless device work of this submit wrote the range
      Mirror *m = findMirror(s->ib.buffer.bytes, s->ib.buffer.size);
      if(m && !m->deviceLocal)
      {
        const size_t lo = (uintptr_t)s->ib.buffer.bytes - (uintptr_t)m->host + s->ib.offset + (size_t)first * isz;
        if(!m->written.overlaps(lo, lo + (size_t)usedVerts * isz))
          ibHost = (const uint8_t *)s->ib.buffer.bytes + s->ib.offset + (size_t)first * isz;
      }
    }
  }
  else if(vertexBound != 0xffffffffu && (uint64_t)first + usedVerts > vertexBound)
    return setError(VB200_ERR_INVALID, "draw reads vertices beyond the bound vertex buffers");

  // ---- targets
  const uint32_t W = s->color.width, H = s->color.height;
  uint8_t *colorDev, *depthDev = nullptr;
  if((rc = resolve(s->color.pixels, (size_t)W * H * 4, ACC_READ | ACC_WRITE | ACC_KEEP_PENDING, &colorDev)))
    return rc;
  const bool depthTest = hasDepth && pl->depth_compare_op != 7u;
  const bool depthWrite = hasDepth && pl->depth_write_enable;
  if(hasDepth && (depthTest || depthWrite))
    if((rc = resolve(s->depth.pixels, (size_t)W * H * 4, ACC_READ | ACC_WRITE | ACC_KEEP_PENDING, &depthDev)))
      return rc;
  if(g.mirrorMerges != mergesBefore && !resolvedTwice && !getenv("VB200_DEBUG_NO_RERESOLVE"))
  {
    resolvedTwice = true;
    goto resolveRanges;
  }

  // Raster back end. A pass is order-independent ("resolvable") unless it blends or runs
  // NOT_EQUAL against a depth buffer it also writes; see scaffold.cu.
  int resolveMode = -1;
  if(!pl->blend_enable)
  {
    if(!depthTest)
      resolveMode = 4;
    else if(!depthWrite)
      resolveMode = 4;    // static test
    else
      switch(pl->depth_compare_op)
      {
        case 0: resolveMode = 4; break;    // NEVER: nothing passes
        case 1: resolveMode = 0; break;    // LESS
        case 2: resolveMode = 4; break;    // EQUAL + write keeps the buffer's value: static test
        case 3: resolveMode = 1; break;    // LESS_OR_EQUAL
        case 4: resolveMode = 2; break;    // GREATER
        case 6: resolveMode = 3; break;    // GREATER_OR_EQUAL
        default: resolveMode = -1; break;  // NOT_EQUAL + write is order dependent
      }
  }
  if(g.optRasterPath == 1)
    resolveMode = -1;
  // a fragment shader that may discard (OpKill, extended mode): the colour a pixel ends up with is that of the
  // last fragment that passed AND was kept, which only the in-order kernel knows
  if(pl->fs->e.uses_kill)
    resolveMode = -1;

  // ---- the batch this draw joins (kernels.h: Vb200BatchDraw). Everything that is uniform across a kernel
  // launch is part of the key; a draw with another key ends the recorded batch and starts the next.
  BatchKey key;
  memset(&key, 0, sizeof(key));
  key.env = env;
  key.vs = pl->vs->serial;
  key.fs = pl->fs->serial;
  key.topology = pl->topology;
  key.frontFace = pl->front_face;
  key.cullMode = pl->cull_mode;
  key.depthOp = pl->depth_compare_op;
  key.depthWrite = depthWrite ? 1u : 0u;
  key.blend[0] = pl->blend_enable ? 1u : 0u;
  key.blend[1] = pl->src_color_blend_factor;
  key.blend[2] = pl->dst_color_blend_factor;
  key.blend[3] = pl->color_blend_op;
  key.colorDev = colorDev;
  key.depthDev = depthDev;
  key.width = W;
  key.height = H;
  key.vertexBound = vertexBound;
  key.nslots = nslots;
  key.tileKernelId = resolveMode >= 0 ? K_TILE_RESOLVE + resolveMode : K_TILE_ORDERED;
  Batch &b = g.batch;
  if(b.active && (b.closed || !g.optBatchDraws || memcmp(&b.key, &key, sizeof(key)) != 0 || b.draws.size() >= 8192 ||
                  b.numTris + (uint64_t)numTris >= (1u << 24) - 1u || b.numVerts + (uint64_t)usedVerts >= (1u << 30)))
    if((rc = flushBatch()))
      return rc;
  if(!b.active)
  {
    b.reset();
    b.key = key;
    if((rc = getKernel(K_VERTEX, pl->vs, &b.kVertex)))
      return rc;
    if((rc = getKernel(key.tileKernelId, pl->fs, &b.kTile, &b.tileThreads)))
      return rc;
    if(!b.tileThreads)
      return setError(VB200_ERR_LINK, "kernel scaffold without a CTA size");
    // deferred clears of exactly these attachments are folded into the tile kernel; anything else that
    // overlaps them is materialised now
    auto take = [&](uint8_t *dev, uint32_t bit, uint32_t *word) {
      auto it = g.pendingClears.find(dev);
      // with sort-first ownership a rank clears only the tiles it owns: colour of the others arrives from
      // their owners (peer stores or the all-gather), their depth is never read on this rank
      if(it != g.pendingClears.end() && it->second.count == (size_t)W * H)
      {
        b.clearFlags |= bit;
        *word = it->second.value;
        g.pendingClears.erase(it);
      }
    };
    uint32_t depthBits = 0;
    take(colorDev, 1u, &b.clearColorWord);
    if(depthDev)
      take(depthDev, 2u, &depthBits);
    memcpy(&b.clearDepthValue, &depthBits, 4);
    if((rc = materializeClears(colorDev, (size_t)W * H * 4)))
      return rc;
    if(depthDev && (rc = materializeClears(depthDev, (size_t)W * H * 4)))
      return rc;
    b.active = true;
  }

  // ---- the draw's vertex span
  Batch::Draw d;
  memset(&d, 0, sizeof(d));
  d.dev.ib = ibDev;
  d.dev.index_type = s->ib.index_type;
  d.dev.first = first;
  d.dev.num_tris = numTris;
  d.dev.tri_base = (uint32_t)b.numTris;
  d.usedVerts = usedVerts;
  if(!indexed)
  {
    d.kind = Batch::SPAN_HOST;
    d.spanBase = first;
    d.spanCount = usedVerts;
  }
  else if(vertexBound == 0xffffffffu)
  {
    // no strided attribute bounds the vertex count (rare: index-only shaders): measure the range now and
    // read it back
    if((rc = flushBatchKeepOpen()))
      return rc;
    g.stats.kernel_launches += vb200::launch_index_range(ibDev, s->ib.index_type, first, usedVerts, g.range, g.stream);
    uint32_t r2[2];
    CU(cudaMemcpyAsync(r2, g.range, 8, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    if(r2[1] < r2[0])
      return VB200_OK;
    d.kind = Batch::SPAN_HOST;
    d.spanBase = r2[0];
    d.spanCount = r2[1] - r2[0] + 1u;
  }
  else if(vertexBound <= usedVerts)
    d.kind = Batch::SPAN_ALL;    // a mesh: every bound vertex is referenced several times, all are shaded
  else if(ibHost && usedVerts <= (1u << 16))
  {
    // a small draw out of a larger vertex buffer, indices readable on the host: [min, max] is measured there
    // when the batch runs, and only if the batch's draws do not reference the whole buffer anyway
    d.kind = Batch::SPAN_INDEXED;
    d.ibHost = ibHost;
  }
  else
    d.kind = Batch::SPAN_DEVICE;    // measured by k_index_range when the batch runs
  b.draws.push_back(d);
  b.numTris += numTris;
  b.numVerts += usedVerts;
  return VB200_OK;
}

int vb200_sample(const vb200_image *tex, int cube, uint64_t byte_offset, const float *uvw, float *out_rgba,
                 size_t count)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(!tex || !tex->pixels || !uvw || !out_rgba)
    return setError(VB200_ERR_INVALID, "sample: NULL argument");
  if(count == 0)
    return VB200_OK;
  uint8_t *dev;
  uint64_t bytes = cube ? sliceBytes(*tex) * 6 : imageBytes(*tex);
  if((rc = resolve(tex->pixels, sampledExtent(tex->pixels, bytes), ACC_READ, &dev)))
    return rc;
  Vb200Image d;
  d.pixels = dev;
  d.width = tex->width;
  d.height = tex->height;
  d.bpp = tex->bytes_per_pixel;
  d.format = tex->format;
  d.mips = tex->mip_levels;
  d.layers = tex->array_layers;
  d.slice_bytes = sliceBytes(*tex);
  const size_t nin = count * (cube ? 3 : 2);
  float *din = nullptr;
  float4 *dout = nullptr;
  CU(cudaMalloc((void **)&din, nin * sizeof(float)));
  CU(cudaMalloc((void **)&dout, count * sizeof(float4)));
  CU(cudaMemcpyAsync(din, uvw, nin * sizeof(float), cudaMemcpyHostToDevice, g.stream));
  g.stats.kernel_launches += vb200::launch_sample(d, cube, byte_offset, din, dout, count, g.stream);
  CU(cudaMemcpyAsync(out_rgba, dout, count * sizeof(float4), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  cudaFree(din);
  cudaFree(dout);
  // synchronous call: it is its own "submit", so host memory becomes authoritative again afterwards
  return vb200_flush();
}

// ---- multi-GPU ---------------------------------------------------------------------------------
int vb200_set_tile_owner(int rank, int world)
{
  if(g.ready)
    flushBatch();
  if(world < 1 || rank < 0 || rank >= world)
    return setError(VB200_ERR_INVALID, "bad tile owner %d/%d", rank, world);
  g.ownerRank = (uint32_t)rank;
  g.ownerWorld = (uint32_t)world;
  return VB200_OK;
}

int vb200_set_peer_targets(const void *local_color_device, void *const *peer_color_device, int num_peers)
{
  if(g.ready)
    flushBatch();
  if(!local_color_device || num_peers < 0 || num_peers > 7 || (num_peers && !peer_color_device))
    return setError(VB200_ERR_INVALID, "set_peer_targets: bad arguments (at most 7 peers)");
  uint8_t *local = (uint8_t *)local_color_device;
  if(num_peers == 0)
  {
    auto it = g.exchange.find(local);
    if(it != g.exchange.end())
    {
      it->second.peers.clear();
      if(!it->second.multicast)
        g.exchange.erase(it);
    }
    return VB200_OK;
  }
  std::vector<uint8_t *> v;
  for(int i = 0; i < num_peers; i++)
  {
    if(!peer_color_device[i])
      return setError(VB200_ERR_INVALID, "set_peer_targets: NULL peer pointer");
    v.push_back((uint8_t *)peer_color_device[i]);
  }
  Context::ExchangeRange &r = g.exchange[local];
  r.exact = true;
  r.peers = v;
  return VB200_OK;
}

int vb200_set_multicast_target(const void *local_color_device, void *multicast_device)
{
  if(g.ready)
    flushBatch();
  if(!local_color_device)
    return setError(VB200_ERR_INVALID, "set_multicast_target: NULL image");
  uint8_t *local = (uint8_t *)local_color_device;
  if(!multicast_device)
  {
    auto it = g.exchange.find(local);
    if(it != g.exchange.end())
    {
      it->second.multicast = nullptr;
      if(it->second.peers.empty())
        g.exchange.erase(it);
    }
    return VB200_OK;
  }
  Context::ExchangeRange &r = g.exchange[local];
  r.exact = true;
  r.multicast = (uint8_t *)multicast_device;
  return VB200_OK;
}

uint32_t vb200_tiles_per_rank(uint32_t width, uint32_t height, int world)
{
  const uint32_t tiles = ((width + VB200_TILE - 1) / VB200_TILE) * ((height + VB200_TILE - 1) / VB200_TILE);
  return world > 0 ? (tiles + (uint32_t)world - 1) / (uint32_t)world : tiles;
}

int vb200_tiles_pack(const vb200_image *image, void *dst_device, uint64_t dst_size)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if((rc = checkTarget(image, "tiles_pack")))
    return rc;
  const uint32_t slots = vb200_tiles_per_rank(image->width, image->height, (int)g.ownerWorld);
  if(dst_size < (uint64_t)slots * 4096)
    return setError(VB200_ERR_INVALID, "tiles_pack: destination too small");
  if(!isDevicePointer(dst_device))
    return setError(VB200_ERR_INVALID, "tiles_pack: destination must be device memory");
  uint8_t *dev;
  if((rc = resolve(image->pixels, (size_t)image->width * image->height * 4, ACC_READ, &dev)))
    return rc;
  g.stats.kernel_launches += vb200::launch_tiles_pack((const uint32_t *)dev, image->width, image->height, g.ownerRank,
                                                      g.ownerWorld, (uint32_t *)dst_device, g.stream);
  CU(cudaGetLastError());
  return VB200_OK;
}

int vb200_tiles_unpack(const vb200_image *image, const void *src_device, uint64_t src_size, int world)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if((rc = checkTarget(image, "tiles_unpack")))
    return rc;
  if(world < 1)
    return setError(VB200_ERR_INVALID, "tiles_unpack: bad world size");
  const uint32_t slots = vb200_tiles_per_rank(image->width, image->height, world);
  if(src_size < (uint64_t)slots * 4096 * world)
    return setError(VB200_ERR_INVALID, "tiles_unpack: source too small");
  if(!isDevicePointer(src_device))
    return setError(VB200_ERR_INVALID, "tiles_unpack: source must be device memory");
  uint8_t *dev;
  if((rc = resolve(image->pixels, (size_t)image->width * image->height * 4, ACC_READ | ACC_WRITE, &dev)))
    return rc;
  g.stats.kernel_launches += vb200::launch_tiles_unpack((uint32_t *)dev, image->width, image->height, (uint32_t)world,
                                                        slots, (const uint32_t *)src_device, g.stream);
  CU(cudaGetLastError());
  return VB200_OK;
}

// ---- introspection -----------------------------------------------------------------------------
int vb200_get_stats(vb200_stats *out)
{
  if(!out)
    return setError(VB200_ERR_INVALID, "get_stats: NULL");
  int rc = requireReady();
  if(rc)
    return rc;
  static Vb200DrawCounters c;
  CU(cudaMemcpyAsync(&c, g.counters, sizeof(c), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  *out = g.stats;
  out->triangles_out = out->fragments_covered = out->fragments_shaded = out->tile_pairs = 0;
  for(int i = 0; i < VB200_COUNTER_SLOTS; i++)
  {
    out->tile_pairs += c.slot[i].tile_pairs;
    out->triangles_out += c.slot[i].triangles_out;
    out->fragments_covered += c.slot[i].fragments_covered;
    out->fragments_shaded += c.slot[i].fragments_shaded;
  }
  return VB200_OK;
}

void vb200_reset_stats(void)
{
  if(g.ready)
    flushBatch();
  memset(&g.stats, 0, sizeof(g.stats));
  if(g.ready)
    cudaMemsetAsync(g.counters, 0, sizeof(Vb200DrawCounters), g.stream);
}

int vb200_get_phase_times(double *ms, uint64_t *counts, int n)
{
  for(int i = 0; i < n && i < VB200_PHASES; i++)
  {
    if(ms)
      ms[i] = g.phaseMs[i];
    if(counts)
      counts[i] = g.phaseCount[i];
  }
  return VB200_OK;
}

int vb200_l2_flush(void)
{
  // benchmark hygiene: overwrite a buffer larger than the 126 MB L2 so the next step starts cold
  int rc = requireReady();
  if(rc)
    return rc;
  static void *scratch = nullptr;
  const size_t bytes = (size_t)512 << 20;
  if(!scratch)
    CU(cudaMalloc(&scratch, bytes));
  static int v = 0;
  CU(cudaMemsetAsync(scratch, ++v & 0xff, bytes, g.stream));
  return VB200_OK;
}

int vb200_event_record(int slot)
{
  int rc = requireReady();
  if(rc)
    return rc;
  if(slot < 0 || slot >= 16)
    return setError(VB200_ERR_INVALID, "event slot out of range");
  if(!g.userEvents[slot])
    CU(cudaEventCreate(&g.userEvents[slot]));
  CU(cudaEventRecord(g.userEvents[slot], g.stream));
  return VB200_OK;
}

int vb200_event_elapsed_ms(int start_slot, int end_slot, float *ms)
{
  if(start_slot < 0 || start_slot >= 16 || end_slot < 0 || end_slot >= 16 || !ms || !g.userEvents[start_slot] ||
     !g.userEvents[end_slot])
    return setError(VB200_ERR_INVALID, "event_elapsed: bad slots");
  CU(cudaEventSynchronize(g.userEvents[end_slot]));
  CU(cudaEventElapsedTime(ms, g.userEvents[start_slot], g.userEvents[end_slot]));
  return VB200_OK;
}

int vb200_set_option(const char *name, int64_t value)
{
  if(!name)
    return setError(VB200_ERR_INVALID, "set_option: NULL name");
  if(g.ready)
    flushBatch();    // recorded draws run under the settings they were recorded with
  if(!strcmp(name, "raster_path"))
    g.optRasterPath = value;
  else if(!strcmp(name, "count_fragments"))
    g.optCountFragments = value;
  else if(!strcmp(name, "fuse_clears"))
    g.optFuseClears = value;
  else if(!strcmp(name, "batch_draws"))
    g.optBatchDraws = value;
  else if(!strcmp(name, "slot_keys"))
    g.optSlotKeys = value;
  else if(!strcmp(name, "tile_list_cap"))
    g.optTileListCap = value;
  else if(!strcmp(name, "mgpu_mirrors"))
    g.optMgpuMirrors = value;
  else if(!strcmp(name, "extended_spirv"))
    vb200::set_extended_spirv(value != 0);
  else if(!strcmp(name, "time_kernels"))
  {
    g.optTimeKernels = value;
    for(int i = 0; i < VB200_PHASES; i++)
    {
      g.phaseMs[i] = 0.0;
      g.phaseCount[i] = 0;
    }
  }
  else
    return setError(VB200_ERR_INVALID, "unknown option '%s'", name);
  return VB200_OK;
}
}    // extern "C"
