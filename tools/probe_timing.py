#!/usr/bin/env python
"""Developer aid (runs on the GPU box): where does a C3 frame's device time go — kernels, the L2 flush's
write-back, or host enqueue gaps?  usage: python tools/probe_timing.py [workload]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from harness import abi, scenes
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
scene = bench.build_scene(wl)
gpu = abi.backend("vb200", 0)
L = gpu.lib
L.vb200_event_elapsed_ms.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
L.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
L.vb200_mem_upload.argtypes = [C.c_void_p, C.c_uint64]
bound = scenes.BoundScene(gpu, scene)
bufs = bench.scene_host_buffers(bound)
for a, _ in bufs:
    L.vb200_mem_register(a.ctypes.data, a.nbytes)
L.vb200_set_sync_mode(1)
for a, is_in in bufs:
    if is_in:
        L.vb200_mem_upload(a.ctypes.data, a.nbytes)
for _ in range(20):
    bound.submit()
gpu.flush()
ms = C.c_float()

def timed(n, flush, sync_each):
    ts = []
    for _ in range(n):
        if flush:
            L.vb200_l2_flush()
        L.vb200_event_record(0)
        bound.submit()
        L.vb200_event_record(1)
        if sync_each:
            L.vb200_event_elapsed_ms(0, 1, C.byref(ms))
            ts.append(ms.value)
    gpu.flush()
    return ts

print("flush+sync each   :", np.mean(timed(20, True, True)))
print("no flush, sync each:", np.mean(timed(20, False, True)))
# back-to-back, no sync: total device time / frames
L.vb200_event_record(2)
t0 = time.perf_counter()
for _ in range(50):
    bound.submit()
t1 = time.perf_counter()
L.vb200_event_record(3)
L.vb200_event_elapsed_ms(2, 3, C.byref(ms))
print("back-to-back 50 frames: device ms/frame", ms.value / 50, " host enqueue ms/frame", (t1 - t0) * 1e3 / 50)
# host time of one submit on an idle GPU
gpu.flush()
t0 = time.perf_counter(); bound.submit(); t1 = time.perf_counter(); gpu.flush(); t2 = time.perf_counter()
print("idle GPU: submit returns after %.3f ms, flush after %.3f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))

# sort-first on ONE GPU: this rank's share of the tile work without any exchange (tail / balance check)
if len(sys.argv) > 2:
    L.vb200_set_tile_owner.argtypes = [C.c_int, C.c_int]
    L.vb200_get_phase_times.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.vb200_set_option.argtypes = [C.c_char_p, C.c_int64]
    world = int(sys.argv[2])
    for r in range(world):
        L.vb200_set_tile_owner(r, world)
        for _ in range(3):
            bound.submit()
        gpu.flush()
        L.vb200_set_option(b"time_kernels", 1)
        gpu.reset_stats()
        for _ in range(5):
            L.vb200_l2_flush()
            bound.submit()
        gpu.flush()
        pm = (C.c_double * 5)(); pc = (C.c_uint64 * 5)()
        L.vb200_get_phase_times(pm, pc, 5)
        L.vb200_set_option(b"time_kernels", 0)
        print("world", world, "rank", r, {n: round(pm[i] / max(pc[i], 1), 4) for i, n in enumerate(["clear", "vertex", "setup", "bin", "tiles"])})
