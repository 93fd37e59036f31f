"""One rank of the sort-first multi-GPU tests (tests/test_mgpu.py starts `world` of these, one per GPU).
usage: python tests/mgpu_worker.py <rank> <world> <session> <mode>
Everything on the data path goes through the C-ABI of libvisor_b200.so: no torch, no NCCL."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harness import abi, scenes  # noqa: E402


def api(gpu):
    L = gpu.lib
    L.vb200_mgpu_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.vb200_mgpu_info.argtypes = [C.POINTER(C.c_int)] * 3
    L.vb200_mgpu_alloc.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
    L.vb200_mgpu_free.argtypes = [C.c_void_p]
    L.vb200_mgpu_push.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mgpu_upload.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_set_option.argtypes = [C.c_char_p, C.c_int64]
    L.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mem_unregister.argtypes = [C.c_void_p]
    L.vb200_mem_upload.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_present.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    return L


def scene_list():
    return [scenes.c3_mesh(640, 360, 160, 80), scenes.c5_textured(333, 211, 100, 50, tex_size=64),
            scenes.c4_particles(320, 200, 3000), scenes.random_triangles(200, 120, 60, 5, max_size=1.2),
            scenes.c1_triangle(97, 65)]


def inputs_of(scene):
    out, seen = [], set()
    for d in scene.draws:
        for a in [v for v, _ in d.vbs] + ([d.ib[0]] if d.ib is not None else []) + [u[2] for u in d.ubos] + \
                [t[2] for t in d.textures]:
            if id(a) not in seen:
                seen.add(id(a))
                out.append(a)
    return out


def main():
    rank, world, session, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    if mode.endswith("-p2p"):
        os.environ["VB200_MGPU_NO_MULTICAST"] = "1"
    gpu = abi.backend("vb200", rank)          # vb200_init(device = rank)
    cpu = abi.backend("vor")
    L = api(gpu)
    gpu.check(L.vb200_mgpu_init(rank, world, rank, session.encode()), "mgpu_init")
    r, w, mc = C.c_int(), C.c_int(), C.c_int()
    L.vb200_mgpu_info(C.byref(r), C.byref(w), C.byref(mc))
    assert (r.value, w.value) == (rank, world)
    print(f"rank {rank}/{world}: multicast={mc.value}", flush=True)
    for sc in scene_list():
        want_c, _ = scenes.render(cpu, sc)
        px = sc.width * sc.height
        if mode.startswith("alloc"):
            # attachments in symmetric buffers, inputs resident in HBM (explicit residency)
            gpu.check(L.vb200_set_sync_mode(1), "set_sync_mode")
            ins = inputs_of(sc)
            for a in ins:
                gpu.check(L.vb200_mem_register(a.ctypes.data, a.nbytes), "mem_register")
                gpu.check(L.vb200_mem_upload(a.ctypes.data, a.nbytes), "mem_upload")
            col, dep = C.c_void_p(), C.c_void_p()
            gpu.check(L.vb200_mgpu_alloc(px * 4, C.byref(col)), "mgpu_alloc")
            gpu.check(L.vb200_mgpu_alloc(px * 4, C.byref(dep)), "mgpu_alloc")
            bound = scenes.BoundScene(gpu, sc, color_device_ptr=col.value, depth_device_ptr=dep.value)
            got = np.zeros((sc.height, sc.width, 4), np.uint8)
            for _ in range(2):    # twice: the second frame runs while nothing is left over from the first
                bound.submit()
                gpu.check(L.vb200_mgpu_barrier(), "mgpu_barrier")
                t = C.c_int()
                gpu.check(L.vb200_present(C.byref(bound.color_img), got.ctypes.data, got.nbytes, C.byref(t)), "present")
                gpu.flush()
                assert np.array_equal(got, want_c), f"rank {rank} {sc.name}: {(got != want_c).any(-1).sum()} pixels differ"
                gpu.check(L.vb200_mgpu_barrier(), "mgpu_barrier")    # nobody clears while a peer still reads
                gpu.flush()
            gpu.check(L.vb200_mgpu_free(col), "mgpu_free")
            gpu.check(L.vb200_mgpu_free(dep), "mgpu_free")
            for a in ins:
                L.vb200_mem_unregister(a.ctypes.data)
            gpu.check(L.vb200_set_sync_mode(0), "set_sync_mode")
        else:
            # "mirrors": host attachments (coherent memory, the ICD's mode); registered ranges get symmetric
            # mirrors, inputs are uploaded in slices, vb200_flush ends with the exchange barrier
            gpu.check(L.vb200_set_option(b"mgpu_mirrors", 1), "set_option")
            bound = scenes.BoundScene(gpu, sc)
            ins = inputs_of(sc)
            regs = ins + [bound.color] + ([bound.depth] if bound.depth is not None else [])
            for a in regs:
                gpu.check(L.vb200_mem_register(a.ctypes.data, a.nbytes), "mem_register")
            gpu.check(L.vb200_set_sync_mode(1), "set_sync_mode")
            for a in ins:
                gpu.check(L.vb200_mgpu_upload(a.ctypes.data, a.nbytes), "mgpu_upload")
            gpu.check(L.vb200_mgpu_barrier(), "mgpu_barrier")
            bound.submit()
            gpu.check(L.vb200_mgpu_barrier(), "mgpu_barrier")
            L.vb200_mem_download.argtypes = [C.c_void_p, C.c_uint64]
            gpu.check(L.vb200_mem_download(bound.color.ctypes.data, bound.color.nbytes), "mem_download")
            gpu.flush()
            assert np.array_equal(bound.color, want_c), f"rank {rank} {sc.name} (sliced upload): image differs"
            # and plainly coherent: the library uploads what the draw reads, flush brings the whole image back
            gpu.check(L.vb200_set_sync_mode(0), "set_sync_mode")
            bound.color[:] = 0
            got_c, _ = bound.run()
            assert np.array_equal(got_c, want_c), f"rank {rank} {sc.name} (coherent): image differs"
            for a in regs:
                L.vb200_mem_unregister(a.ctypes.data)
            gpu.check(L.vb200_set_option(b"mgpu_mirrors", 0), "set_option")
    gpu.check(L.vb200_mgpu_shutdown(), "mgpu_shutdown")
    print(f"rank {rank}: ok", flush=True)


if __name__ == "__main__":
    main()
