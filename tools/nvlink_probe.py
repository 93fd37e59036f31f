#!/usr/bin/env python
"""Developer aid, ONE process on a box with >= 2 GPUs (so that it can run under ncu, which a multi-rank command
cannot): rank 0's share of a sort-first C5 frame with the colour image of a second GPU as peer target. Every
pixel of the owned tiles is stored locally and over NVLink into the peer image (vb200_set_peer_targets, the
unicast flavour of the fused exchange); the peer image is checked against the local one afterwards.
  ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,lts__t_sectors_srcunit_ltcfabric.sum,gpu__time_duration.sum \\
      -k regex:resolve -s 6 -c 2 python tools/nvlink_probe.py [world]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from harness import abi, scenes  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
assert torch.cuda.device_count() >= 2, "needs two GPUs"
gpu = abi.backend("vb200", 0)
L = gpu.lib
L.vb200_set_peer_targets.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int]
L.vb200_mem_device_ptr.restype = C.c_void_p
L.vb200_mem_device_ptr.argtypes = [C.c_void_p]
L.vb200_set_sync_mode.argtypes = [C.c_int]
L.vb200_mem_upload.argtypes = [C.c_void_p, C.c_uint64]
sc = scenes.c5_textured()
W, H = sc.width, sc.height
# colour images: local on cuda:0, the "other rank's" on cuda:1, mapped into this process (peer access)
local = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
peer = torch.zeros((H, W), dtype=torch.int32, device="cuda:1")
with torch.cuda.device(0):
    assert torch.cuda.can_device_access_peer(0, 1)
    torch.zeros(1, device="cuda:0").copy_(torch.zeros(1, device="cuda:1"))    # makes torch enable peer access 0 <-> 1
    rt = C.CDLL("libcudart.so")
    rt.cudaSetDevice(0)
    rt.cudaDeviceEnablePeerAccess(1, 0)    # (already enabled: returns an error code that is fine to ignore)
    rt.cudaGetLastError()
bound = scenes.BoundScene(gpu, sc, color_device_ptr=local.data_ptr())
gpu.check(L.vb200_set_sync_mode(1), "set_sync_mode")    # explicit residency: inputs uploaded once, nothing copied per frame
from bench import scene_host_buffers  # noqa: E402

L.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
for a, is_in in scene_host_buffers(bound):
    if a is bound.color:
        continue    # (the colour attachment is the device image above)
    gpu.check(L.vb200_mem_register(a.ctypes.data, a.nbytes), "mem_register")
    if is_in:
        gpu.check(L.vb200_mem_upload(a.ctypes.data, a.nbytes), "mem_upload")
gpu.check(L.vb200_set_tile_owner(0, world), "set_tile_owner")
ptrs = (C.c_void_p * 1)(peer.data_ptr())
gpu.check(L.vb200_set_peer_targets(local.data_ptr(), ptrs, 1), "set_peer_targets")
for _ in range(8):
    bound.submit()
    gpu.flush()
torch.cuda.synchronize(0)
a, b = local.cpu().numpy(), peer.cpu().numpy()
from harness import tiles  # noqa: E402

own = tiles.owned_mask(W, H, 0, world)
print("owned pixels", int(own.sum()), "bytes over NVLink per frame (expected)", int(own.sum()) * 4,
      "peer image equals local on the owned tiles:", bool(np.array_equal(a[own], b[own])),
      "peer untouched elsewhere:", bool((b[~own] == 0).all()))
