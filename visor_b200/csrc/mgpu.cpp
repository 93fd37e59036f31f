// mgpu.cpp — sort-first rendering across the GPUs of one node, one process per GPU (include/visor_b200.h,
// "sort-first multi-GPU"). No reference equivalent: visor is a single-device CPU renderer.
//
// What lives here is the plumbing the fused exchange of the tile kernels needs, without any framework in the
// data plane:
//   * a bootstrap between the ranks of a session over abstract unix-domain sockets (file descriptors of
//     allocation handles travel as SCM_RIGHTS ancillary data),
//   * symmetric buffers: one cuMemCreate allocation per rank, every rank maps all of them (NVLink peer
//     mappings) and — where the NVSwitch supports it — one multicast object bound to all of them, mapped once
//     more, so that a single multimem.st reaches every rank,
//   * a device-side barrier over flags in such a buffer, and a push kernel that replicates a slice of a
//     symmetric buffer to all ranks (the sharded upload of the frame's inputs).
// The driver API is reached through cudaGetDriverEntryPoint, so the library still loads on a machine
// without a driver.
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/visor_b200.h"
#include "kernels.h"
#include "runtime_internal.h"

namespace
{
using vb200::set_error;
using vb200::Vb200PeerSet;

// ---- driver API ------------------------------------------------------------------------------
struct Driver
{
  bool loaded = false, ok = false;
  CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
  CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long);
  CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags);
  CUresult (*DeviceGet)(CUdevice *, int);
  CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice);
  CUresult (*GetErrorString)(CUresult, const char **);
} drv;

template <typename F>
bool entry(const char *name, F *out)
{
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
  {
    cudaGetLastError();
    return false;
  }
  *out = (F)fn;
  return true;
}

bool loadDriver()
{
  if(drv.loaded)
    return drv.ok;
  drv.loaded = true;
  drv.ok = entry("cuMemGetAllocationGranularity", &drv.MemGetAllocationGranularity) &&
           entry("cuMemCreate", &drv.MemCreate) && entry("cuMemRelease", &drv.MemRelease) &&
           entry("cuMemAddressReserve", &drv.MemAddressReserve) && entry("cuMemAddressFree", &drv.MemAddressFree) &&
           entry("cuMemMap", &drv.MemMap) && entry("cuMemUnmap", &drv.MemUnmap) &&
           entry("cuMemSetAccess", &drv.MemSetAccess) &&
           entry("cuMemExportToShareableHandle", &drv.MemExportToShareableHandle) &&
           entry("cuMemImportFromShareableHandle", &drv.MemImportFromShareableHandle) &&
           entry("cuDeviceGet", &drv.DeviceGet) && entry("cuDeviceGetAttribute", &drv.DeviceGetAttribute) &&
           entry("cuGetErrorString", &drv.GetErrorString);
  if(drv.ok)
  {
    // multicast entry points are optional (older drivers): without them the exchange uses peer stores
    if(!(entry("cuMulticastCreate", &drv.MulticastCreate) && entry("cuMulticastAddDevice", &drv.MulticastAddDevice) &&
         entry("cuMulticastBindMem", &drv.MulticastBindMem) &&
         entry("cuMulticastGetGranularity", &drv.MulticastGetGranularity)))
      drv.MulticastCreate = nullptr;
  }
  return drv.ok;
}

const char *cuErr(CUresult r)
{
  const char *s = nullptr;
  if(drv.GetErrorString)
    drv.GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

#define DRV(call)                                                                                          \
  do                                                                                                       \
  {                                                                                                        \
    CUresult _r = (call);                                                                                  \
    if(_r != CUDA_SUCCESS)                                                                                 \
      return set_error(VB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cuErr(_r), __FILE__, __LINE__);      \
  } while(0)

// ---- state -----------------------------------------------------------------------------------
struct Symmetric
{
  size_t bytes = 0;                      // mapped size (granularity multiple)
  CUmemGenericAllocationHandle local = 0;
  std::vector<CUmemGenericAllocationHandle> imported;    // per rank (own slot unused)
  std::vector<CUdeviceptr> base;         // per rank: where that rank's copy is mapped here (own slot = local copy)
  CUmemGenericAllocationHandle mcHandle = 0;
  CUdeviceptr mc = 0;                    // multicast mapping, 0 if none
};

struct Message
{
  uint32_t magic, seq, rank, has_fd;
};

struct State
{
  bool on = false;
  int rank = 0, world = 1, device = 0;
  std::string session;
  int listenFd = -1;
  uint32_t seq = 0;
  bool multicast = false;
  std::map<std::pair<uint32_t, uint32_t>, int> inbox;    // (seq, rank) -> fd (or -1)
  std::map<uint8_t *, Symmetric> allocs;                   // by local base
  Symmetric signal;
  uint32_t epoch = 0;
  uint32_t *timedOut = nullptr;          // mapped pinned word the barrier kernel raises
  uint32_t *timedOutDev = nullptr;
} M;

// ---- bootstrap -------------------------------------------------------------------------------
void socketName(int rank, sockaddr_un *addr, socklen_t *len)
{
  memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  // abstract namespace (leading NUL): no file to clean up, vanishes with the process
  const int n = snprintf(addr->sun_path + 1, sizeof(addr->sun_path) - 1, "visor_b200.%s.%d", M.session.c_str(), rank);
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

int sendTo(int peer, const Message &msg, int fd)
{
  sockaddr_un addr;
  socklen_t alen;
  socketName(peer, &addr, &alen);
  int s = -1;
  for(int attempt = 0; attempt < 3000; attempt++)    // the peer may not be listening yet: up to ~60 s
  {
    s = socket(AF_UNIX, SOCK_STREAM, 0);
    if(s < 0)
      return set_error(VB200_ERR_CUDA, "mgpu: socket(): %s", strerror(errno));
    if(connect(s, (sockaddr *)&addr, alen) == 0)
      break;
    close(s);
    s = -1;
    timespec ts = {0, 20 * 1000 * 1000};
    nanosleep(&ts, nullptr);
  }
  if(s < 0)
    return set_error(VB200_ERR_CUDA, "mgpu: rank %d of session '%s' is not reachable", peer, M.session.c_str());
  iovec iov = {(void *)&msg, sizeof(msg)};
  msghdr mh;
  memset(&mh, 0, sizeof(mh));
  mh.msg_iov = &iov;
  mh.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))];
  if(fd >= 0)
  {
    memset(ctrl, 0, sizeof(ctrl));
    mh.msg_control = ctrl;
    mh.msg_controllen = sizeof(ctrl);
    cmsghdr *c = CMSG_FIRSTHDR(&mh);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
  }
  const ssize_t n = sendmsg(s, &mh, 0);
  close(s);
  if(n != (ssize_t)sizeof(msg))
    return set_error(VB200_ERR_CUDA, "mgpu: sendmsg to rank %d: %s", peer, strerror(errno));
  return VB200_OK;
}

// receive one message from whoever connects next into the inbox
int receiveOne()
{
  // accept with a deadline so that a dead peer surfaces as an error instead of a hang
  timeval tv = {60, 0};
  setsockopt(M.listenFd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  const int s = accept(M.listenFd, nullptr, nullptr);
  if(s < 0)
    return set_error(VB200_ERR_CUDA, "mgpu: no message from the other ranks (%s)", strerror(errno));
  Message msg;
  iovec iov = {&msg, sizeof(msg)};
  msghdr mh;
  memset(&mh, 0, sizeof(mh));
  mh.msg_iov = &iov;
  mh.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))];
  mh.msg_control = ctrl;
  mh.msg_controllen = sizeof(ctrl);
  const ssize_t n = recvmsg(s, &mh, MSG_WAITALL);
  close(s);
  if(n != (ssize_t)sizeof(msg) || msg.magic != 0x56423230u)
    return set_error(VB200_ERR_CUDA, "mgpu: malformed bootstrap message");
  int fd = -1;
  for(cmsghdr *c = CMSG_FIRSTHDR(&mh); c; c = CMSG_NXTHDR(&mh, c))
    if(c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS)
      memcpy(&fd, CMSG_DATA(c), sizeof(int));
  M.inbox[{msg.seq, msg.rank}] = fd;
  return VB200_OK;
}

// Collective: every rank contributes one file descriptor (or -1) and receives everybody else's.
// `only_from` >= 0: only that rank sends (a broadcast); the others just receive from it.
int exchangeFds(int myFd, std::vector<int> &fds, int only_from = -1)
{
  const uint32_t seq = ++M.seq;
  fds.assign(M.world, -1);
  Message msg = {0x56423230u, seq, (uint32_t)M.rank, myFd >= 0 ? 1u : 0u};
  if(only_from < 0 || only_from == M.rank)
    for(int p = 0; p < M.world; p++)
      if(p != M.rank)
        if(int rc = sendTo(p, msg, myFd))
          return rc;
  for(int p = 0; p < M.world; p++)
  {
    if(p == M.rank || (only_from >= 0 && p != only_from))
      continue;
    while(M.inbox.find({seq, (uint32_t)p}) == M.inbox.end())
      if(int rc = receiveOne())
        return rc;
    fds[p] = M.inbox[{seq, (uint32_t)p}];
    M.inbox.erase({seq, (uint32_t)p});
  }
  return VB200_OK;
}

int hostBarrier()
{
  std::vector<int> fds;
  return exchangeFds(-1, fds);
}

// ---- symmetric buffers -----------------------------------------------------------------------
int mapHandle(CUmemGenericAllocationHandle h, size_t bytes, size_t gran, CUdeviceptr *va)
{
  DRV(drv.MemAddressReserve(va, bytes, gran, 0, 0));
  DRV(drv.MemMap(*va, bytes, 0, h, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = M.device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  DRV(drv.MemSetAccess(*va, bytes, &acc, 1));
  return VB200_OK;
}

int symCreate(size_t want, Symmetric &out)
{
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = M.device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  DRV(drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  CUmulticastObjectProp mcp;
  memset(&mcp, 0, sizeof(mcp));
  if(M.multicast)
  {
    mcp.numDevices = (unsigned)M.world;
    mcp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mcp.size = want;
    size_t mg = 0;
    DRV(drv.MulticastGetGranularity(&mg, &mcp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    gran = gran > mg ? gran : mg;
  }
  const size_t bytes = (want + gran - 1) / gran * gran;
  out = Symmetric();
  out.bytes = bytes;
  out.imported.assign(M.world, 0);
  out.base.assign(M.world, 0);
  DRV(drv.MemCreate(&out.local, bytes, &prop, 0));
  int myFd = -1;
  DRV(drv.MemExportToShareableHandle(&myFd, out.local, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  std::vector<int> fds;
  int rc = exchangeFds(myFd, fds);
  close(myFd);
  if(rc)
    return rc;
  for(int p = 0; p < M.world; p++)
  {
    CUmemGenericAllocationHandle h = out.local;
    if(p != M.rank)
    {
      DRV(drv.MemImportFromShareableHandle(&out.imported[p], (void *)(uintptr_t)fds[p],
                                           CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      close(fds[p]);
      h = out.imported[p];
    }
    if((rc = mapHandle(h, bytes, gran, &out.base[p])))
      return rc;
  }
  if(M.multicast)
  {
    mcp.size = bytes;
    int mcFd = -1;
    if(M.rank == 0)
    {
      DRV(drv.MulticastCreate(&out.mcHandle, &mcp));
      DRV(drv.MemExportToShareableHandle(&mcFd, out.mcHandle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    }
    rc = exchangeFds(mcFd, fds, 0);
    if(mcFd >= 0)
      close(mcFd);
    if(rc)
      return rc;
    if(M.rank != 0)
    {
      DRV(drv.MemImportFromShareableHandle(&out.mcHandle, (void *)(uintptr_t)fds[0],
                                           CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      close(fds[0]);
    }
    CUdevice dev;
    DRV(drv.DeviceGet(&dev, M.device));
    DRV(drv.MulticastAddDevice(out.mcHandle, dev));
    if((rc = hostBarrier()))    // memory may only be bound once every device has been added
      return rc;
    DRV(drv.MulticastBindMem(out.mcHandle, 0, out.local, 0, bytes, 0));
    if((rc = hostBarrier()))
      return rc;
    if((rc = mapHandle(out.mcHandle, bytes, gran, &out.mc)))
      return rc;
  }
  // the allocation starts zeroed on every rank before anybody can store into it
  if(cudaMemsetAsync((void *)out.base[M.rank], 0, bytes, vb200::library_stream()) != cudaSuccess ||
     cudaStreamSynchronize(vb200::library_stream()) != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "mgpu: clearing a symmetric buffer failed");
  return hostBarrier();
}

void symDestroy(Symmetric &s)
{
  if(s.mc)
  {
    drv.MemUnmap(s.mc, s.bytes);
    drv.MemAddressFree(s.mc, s.bytes);
  }
  for(size_t p = 0; p < s.base.size(); p++)
    if(s.base[p])
    {
      drv.MemUnmap(s.base[p], s.bytes);
      drv.MemAddressFree(s.base[p], s.bytes);
    }
  for(CUmemGenericAllocationHandle h : s.imported)
    if(h)
      drv.MemRelease(h);
  if(s.mcHandle)
    drv.MemRelease(s.mcHandle);
  if(s.local)
    drv.MemRelease(s.local);
  s = Symmetric();
}

Vb200PeerSet peerSet(const Symmetric &s, size_t offset = 0)
{
  Vb200PeerSet ps;
  memset(&ps, 0, sizeof(ps));
  for(int p = 0; p < M.world && p < 8; p++)
    ps.peer[p] = (void *)(s.base[p] + offset);
  return ps;
}

Symmetric *findAlloc(const void *dev, size_t *offset)
{
  auto it = M.allocs.upper_bound((uint8_t *)dev);
  if(it == M.allocs.begin())
    return nullptr;
  --it;
  if((uint8_t *)dev >= it->first + it->second.bytes)
    return nullptr;
  if(offset)
    *offset = (size_t)((uint8_t *)dev - it->first);
  return &it->second;
}

int requireOn()
{
  if(!M.on)
    return set_error(VB200_ERR_INVALID, "multi-GPU mode is not initialised (vb200_mgpu_init)");
  return VB200_OK;
}
}    // namespace

namespace vb200
{
bool mgpu_active()
{
  return M.on;
}

bool mgpu_is_symmetric(const uint8_t *dev)
{
  return M.on && findAlloc(dev, nullptr) != nullptr;
}

int mgpu_sym_alloc(size_t bytes, uint8_t **local)
{
  *local = nullptr;
  if(int rc = requireOn())
    return rc;
  Symmetric s;
  if(int rc = symCreate(bytes, s))
    return rc;
  uint8_t *base = (uint8_t *)s.base[M.rank];
  std::vector<uint8_t *> peers;
  for(int p = 0; p < M.world; p++)
    if(p != M.rank)
      peers.push_back((uint8_t *)s.base[p]);
  set_exchange_range(base, s.bytes, peers, (uint8_t *)s.mc);
  M.allocs[base] = s;
  *local = base;
  return VB200_OK;
}

int mgpu_sym_free(uint8_t *local)
{
  auto it = M.allocs.find(local);
  if(it == M.allocs.end())
    return set_error(VB200_ERR_INVALID, "mgpu: %p is not the base of a symmetric buffer", (void *)local);
  cudaStreamSynchronize(library_stream());
  int rc = hostBarrier();    // nobody may still be storing into it
  clear_exchange_range(local);
  symDestroy(it->second);
  M.allocs.erase(it);
  return rc;
}
}    // namespace vb200

// =================================================================================================
extern "C" {

int vb200_mgpu_init(int rank, int world, int device, const char *session)
{
  if(M.on)
    return set_error(VB200_ERR_INVALID, "mgpu: already initialised");
  if(world < 1 || world > 8 || rank < 0 || rank >= world || !session || !*session || strlen(session) > 60)
    return set_error(VB200_ERR_INVALID, "mgpu_init: rank %d / world %d (at most 8) / session", rank, world);
  int rc = vb200_init(device);
  if(rc)
    return rc;
  if(vb200::library_device() != device)
    return set_error(VB200_ERR_INVALID, "mgpu_init: the library is already running on device %d", vb200::library_device());
  if(!loadDriver())
    return set_error(VB200_ERR_CUDA, "mgpu: the CUDA driver's virtual memory management entry points are unavailable");
  M.rank = rank;
  M.world = world;
  M.device = device;
  M.session = session;
  M.seq = 0;
  sockaddr_un addr;
  socklen_t alen;
  socketName(rank, &addr, &alen);
  M.listenFd = socket(AF_UNIX, SOCK_STREAM, 0);
  if(M.listenFd < 0 || bind(M.listenFd, (sockaddr *)&addr, alen) != 0 || listen(M.listenFd, 64) != 0)
  {
    if(M.listenFd >= 0)
      close(M.listenFd);
    M.listenFd = -1;
    return set_error(VB200_ERR_CUDA, "mgpu: cannot listen on the session socket (%s); is the session name in use?",
                     strerror(errno));
  }
  // NVSwitch multicast: only if every entry point exists and the device reports support
  M.multicast = false;
  if(drv.MulticastCreate && world > 1 && !getenv("VB200_MGPU_NO_MULTICAST"))
  {
    CUdevice dev;
    int supported = 0;
    if(drv.DeviceGet(&dev, device) == CUDA_SUCCESS &&
       drv.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS)
      M.multicast = supported != 0;
  }
  M.on = true;
  auto fail = [&](int code) {
    vb200_mgpu_shutdown();
    return code;
  };
  if((rc = hostBarrier()))
    return fail(rc);
  if((rc = symCreate(4096, M.signal)))
    return fail(rc);
  if(cudaHostAlloc((void **)&M.timedOut, sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
     cudaHostGetDevicePointer((void **)&M.timedOutDev, M.timedOut, 0) != cudaSuccess)
    return fail(set_error(VB200_ERR_CUDA, "mgpu: cannot allocate the barrier's status word"));
  *M.timedOut = 0;
  M.epoch = 0;
  if((rc = vb200_set_tile_owner(rank, world)))
    return fail(rc);
  return VB200_OK;
}

int vb200_mgpu_shutdown(void)
{
  if(!M.on)
    return VB200_OK;
  if(vb200::library_stream())
    cudaStreamSynchronize(vb200::library_stream());
  for(auto &kv : M.allocs)
  {
    vb200::clear_exchange_range(kv.first);
    symDestroy(kv.second);
  }
  M.allocs.clear();
  symDestroy(M.signal);
  if(M.timedOut)
    cudaFreeHost(M.timedOut);
  M.timedOut = M.timedOutDev = nullptr;
  for(auto &kv : M.inbox)
    if(kv.second >= 0)
      close(kv.second);
  M.inbox.clear();
  if(M.listenFd >= 0)
    close(M.listenFd);
  M.listenFd = -1;
  M.on = false;
  vb200_set_tile_owner(0, 1);
  return VB200_OK;
}

int vb200_mgpu_info(int *rank, int *world, int *multicast)
{
  if(rank)
    *rank = M.on ? M.rank : 0;
  if(world)
    *world = M.on ? M.world : 1;
  if(multicast)
    *multicast = M.on && M.multicast ? 1 : 0;
  return VB200_OK;
}

int vb200_mgpu_alloc(uint64_t bytes, void **local_device)
{
  if(!local_device || !bytes)
    return set_error(VB200_ERR_INVALID, "mgpu_alloc: empty request");
  uint8_t *p = nullptr;
  int rc = vb200::mgpu_sym_alloc((size_t)bytes, &p);
  *local_device = p;
  return rc;
}

int vb200_mgpu_free(void *local_device)
{
  if(int rc = requireOn())
    return rc;
  return vb200::mgpu_sym_free((uint8_t *)local_device);
}

int vb200_mgpu_barrier(void)
{
  if(int rc = requireOn())
    return rc;
  if(*M.timedOut)
    return set_error(VB200_ERR_CUDA, "mgpu: an earlier cross-rank barrier timed out (a rank is missing)");
  vb200::count_launches(vb200::launch_mgpu_barrier(peerSet(M.signal), (uint32_t)M.rank, (uint32_t)M.world, ++M.epoch,
                                                   M.timedOutDev, vb200::library_stream()));
  if(cudaGetLastError() != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "mgpu: barrier launch failed");
  return VB200_OK;
}

int vb200_mgpu_push(const void *local_device, uint64_t bytes)
{
  if(int rc = requireOn())
    return rc;
  size_t off = 0;
  Symmetric *s = findAlloc(local_device, &off);
  if(!s || off + bytes > s->bytes)
    return set_error(VB200_ERR_INVALID, "mgpu_push: the range is not inside a symmetric buffer");
  if((off | bytes) & 15)
    return set_error(VB200_ERR_INVALID, "mgpu_push: offset and size must be multiples of 16");
  vb200::count_launches(vb200::launch_mgpu_push(peerSet(*s), (void *)s->mc, (uint32_t)M.rank, (uint32_t)M.world, off,
                                                bytes, vb200::library_stream()));
  if(cudaGetLastError() != cudaSuccess)
    return set_error(VB200_ERR_CUDA, "mgpu: push launch failed");
  return VB200_OK;
}

}    // extern "C"
