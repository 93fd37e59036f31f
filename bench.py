#!/usr/bin/env python
"""bench.py — throughput of visor's draw-execution hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c1..c6]

A "step" is one pass of the hot path over one synthetic frame: ClearTarget(colour), ClearTarget(depth),
DrawTriangles — the reference's own operator sequence for a render pass (cmd_exec.cpp:35-142).

N = 1   workload c3: 1 000 000-triangle indexed lit mesh, depth LESS + write, 3840x2160 (BASELINE.json
        configs[2], the scene the north-star roofline target is stated on).
N > 1   workload c5: 4 000 000-triangle textured mesh at 7680x4320, sort-first over screen tiles
        (tile t on rank t % N), geometry replicated, colour assembled by an NCCL all-gather.

value   Mtri/s with inputs already resident in HBM, timed with CUDA events on the library stream,
        L2 flushed (512 MB write) before every timed step, max over ranks.
e2e     the same metric through the C-ABI with HOST buffers: every step uploads the vertex/index/
        uniform data from pinned host memory and downloads colour + depth (the reference's coherent
        host-visible memory semantics), wall-clock around submit..flush.
roofline  for the dominant kernel (tile raster): algorithmic bytes = 8 B/pixel (colour + depth written
        once) / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference: visor's own rasterizer.cpp + texture_sampling.cpp compiled unmodified
        (oracle/_ref), timed on this box's host cores on the same scene.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from harness import abi, scenes, tiles, vkdriver  # noqa: E402


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index: int) -> None:
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = str(e)

    def _run(self) -> None:
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def nvlink_kib(index: int):
    """(tx, rx) NVLink payload counters of one GPU in KiB, summed over its links (NVML field values), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        tx_id = getattr(pynvml, "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX", 138)
        rx_id = getattr(pynvml, "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX", 139)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(tx_id, 0xffffffff), (rx_id, 0xffffffff)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal))
        return tuple(out)
    except Exception:
        return None


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant (tile) kernel, from the
# `ncu --set full` captures summarised under profiles/ (a profiler cannot run inside the timed bench, so
# the per-launch figure of the same build and workload is recorded here; null when none was captured).
NCU_TRAFFIC_SOURCE = "profiles/r02_ncu_c{2,3,4,5}_kernels.txt (ncu --set full, per launch)"
NCU_TRAFFIC = {
    ("c2", 1, "vb200_k_tile_resolve_min_first"): 406272 + 0,
    ("c3", 1, "vb200_k_tile_resolve_min_first"): 27128000 + 10336000,
    ("c4", 1, "vb200_k_tile_ordered"): 34586000 + 169728,
    ("c5", 1, "vb200_k_tile_resolve_min_first"): 148613000 + 209010000,
}


# warp instructions one launch of the tile kernel executes (ncu smsp__inst_executed.sum, same captures): the
# kernel is bound by instruction issue, not by HBM, so the bench also reports its issue-slot utilisation
NCU_WARP_INSTRUCTIONS = {
    ("c2", 1, "vb200_k_tile_resolve_min_first"): 8.94e6,
    ("c3", 1, "vb200_k_tile_resolve_min_first"): 97.70e6,
    ("c4", 1, "vb200_k_tile_ordered"): 249.22e6,
    ("c5", 1, "vb200_k_tile_resolve_min_first"): 538.96e6,
}


# what binds the tile kernels instead of HBM (same ncu captures, profiles/r02_ncu_*_kernels.txt): share of the peak
# the l1tex data pipe (shared-memory + global wavefronts) and the issue slots are busy, per launch
NCU_SM_LIMITERS = {
    ("c2", 1, "vb200_k_tile_resolve_min_first"): {"l1tex_data_pipe": 0.165, "issue_slots": 0.414},
    ("c3", 1, "vb200_k_tile_resolve_min_first"): {"l1tex_data_pipe": 0.652, "issue_slots": 0.698},
    ("c4", 1, "vb200_k_tile_ordered"): {"l1tex_data_pipe": 0.468, "issue_slots": 0.782},
    ("c5", 1, "vb200_k_tile_resolve_min_first"): {"l1tex_data_pipe": 0.753, "issue_slots": 0.737},
}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_scene(workload: str) -> scenes.Scene:
    return {
        "c1": lambda: scenes.c1_triangle(),
        "c2": lambda: scenes.c2_cube(),
        "c3": lambda: scenes.c3_mesh(),
        "c4": lambda: scenes.c4_particles(),
        "c5": lambda: scenes.c5_textured(),
        "c6": lambda: scenes.c6_many_draws(),
    }[workload]()


WORKLOAD_DESC = {
    "c1": "c1: single vkCmdDraw triangle, passthrough VS/FS, 1280x720",
    "c2": "c2: textured cube, D32 depth LESS, bilinear RGBA8, 1920x1080",
    "c3": "c3: 1M-triangle indexed mesh, per-vertex lighting, depth LESS+write, 3840x2160",
    "c4": "c4: 200k alpha-blended quads (400k triangles), ~8x overdraw, 1920x1080",
    "c5": "c5: 4M-triangle textured lit mesh, depth LESS+write, 7680x4320",
    "c6": "c6 (diagnostic, not a BASELINE config): the c3 mesh recorded as 1000 indexed draws of 1000 triangles",
}


def scene_host_buffers(bound: scenes.BoundScene):
    """Every host array a frame touches: (array, is_input)."""
    seen, out = set(), []
    for _, d in bound.calls:
        arrs = [v for v, _ in d.vbs] + ([d.ib[0]] if d.ib is not None else []) + \
            [u[2] for u in d.ubos] + [t[2] for t in d.textures]
        for a in arrs:
            if id(a) not in seen:
                seen.add(id(a))
                out.append((a, True))
    out.append((bound.color, False))
    if bound.depth is not None:
        out.append((bound.depth, False))
    return out


# ------------------------------------------------------------------------------------------------
def reference_native_shaders(on: bool) -> None:
    """The reference JITs shaders to native code (LLVM 6, not buildable here). oracle/ref/ref_glue.cpp serves
    spirv_compile.h with natively compiled C++ equivalents of the bench scenes' shaders (bit-identical to the
    interpreter, tests/test_oracle_vs_ref.py) or, with 0, with the SPIR-V interpreter for everything."""
    ref = abi.backend("vref", 0)
    fn = ref.lib.vref_native_shaders
    fn.argtypes = [C.c_int]
    fn(1 if on else 0)
    # entries are cached per backend and module: drop the reference's so that the setting takes effect
    scenes._shader_cache.mods = {k: v for k, v in scenes._shader_cache.mods.items() if k[0] != "vref"}


def time_reference(scene: scenes.Scene, steps: int, warmup: int, threaded_steps: int = 1, hash_depth: bool = True,
                   interpreted_frames: int = 0):
    """visor's own CPU rasterizer (oracle/_ref) on this box's host cores: threaded as shipped
    (7 workers + main, rasterizer.cpp:8) and serial (the deterministic parity mode)."""
    if not abi.available("vref"):
        return None
    res = {}
    tris = scene.triangles()
    # threaded must come first: the pool cannot be restarted once shut down (rast.kill stays set)
    ref = abi.backend("vref", 1 if threaded_steps > 0 else 0)
    b = scenes.BoundScene(ref, scene)
    if threaded_steps > 0 and ref.fn("threads")() == 8:
        ts = []
        for _ in range(max(1, threaded_steps)):
            t0 = time.perf_counter()
            b.run()
            ts.append(time.perf_counter() - t0)
        res["threaded"] = {"s_per_frame": min(ts), "mtri_s": tris / min(ts) / 1e6, "threads": 8}
    ref.fn("init")(0)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        b.run()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    res["serial"] = {"s_per_frame": float(np.mean(ts)), "mtri_s": tris / float(np.mean(ts)) / 1e6, "threads": 1}
    res["hash"] = scenes.image_hash(b.color, b.depth if hash_depth else None)
    if interpreted_frames > 0:
        # the same frame with the shader stage interpreted (what round 1 reported as the CPU baseline)
        reference_native_shaders(False)
        bi = scenes.BoundScene(ref, scene)
        ts = []
        for _ in range(interpreted_frames):
            t0 = time.perf_counter()
            bi.run()
            ts.append(time.perf_counter() - t0)
        res["serial_interpreted"] = {"s_per_frame": min(ts), "mtri_s": tris / min(ts) / 1e6, "threads": 1}
        res["interpreted_hash_equal"] = scenes.image_hash(bi.color, bi.depth if hash_depth else None) == res["hash"]
        reference_native_shaders(True)
    return res


def workload_config(workload: str, scene, world: int, parallelism: str, raster_path: str) -> dict:
    """`config` of the JSON line: the same keys for both arms"""
    return {"workload": WORKLOAD_DESC[workload], "triangles": scene.triangles(),
            "resolution": [scene.width, scene.height], "l2": "flushed (512 MB write) before each timed step",
            "parallelism": parallelism, "raster_path": raster_path}


PARALLELISM_DESC = {
    "single": "single GPU",
    "multicast": "sort-first x{n}, fused: one NVSwitch-multicast store per pixel from the tile kernel + barrier",
    "p2p": "sort-first x{n}, fused: NVLink peer stores from the tile kernel + barrier",
    "allgather": "sort-first x{n}, NCCL all-gather of owned tiles",
}
RASTER_PATH_DESC = {
    "auto": "auto: visibility-resolve tiles for order-independent passes, ordered tiles otherwise",
    "ordered": "ordered tiles (forced)",
}


def run_reference_arm(args, workload: str) -> None:
    """The reference's own CPU implementation of the path, timed where SURVEY.md §8d says: host wall clock
    around vkQueueSubmit (+ vkQueueWaitIdle) of the unmodified reference ICD (oracle/_ref/libvisor_ref.so:
    icd_interface / cmd_record / cmd_exec / rasterizer / texture_sampling ... compiled where they lie) replaying
    the frame's command buffer, on this box's host cores. Exactly --steps frames after --warmup frames."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not (abi.available("vref") and vkdriver.available()):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference ICD + Vulkan driver) not built"}))
        return
    scene = build_scene(workload)
    tris = scene.triangles()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    modes = {}
    # threaded as shipped (7 workers + the submitting thread, rasterizer.cpp:8) must run before the workers are
    # joined for the serial mode; one frame of it (it is slower on small-triangle scenes: one mutex-guarded
    # queue, rasterizer.cpp:473-485 — and it races on the framebuffer, so it is never the parity mode)
    if workload != "c5":
        vkdriver.run(vkdriver.ICD_REF, scene, frames=1, serial_reference=False)
        t = vkdriver.frame_seconds(1)[0]
        modes["threaded"] = {"s_per_frame": t, "mtri_s": tris / t / 1e6, "threads": 8, "frames": 1}
    vkdriver.run(vkdriver.ICD_REF, scene, frames=warm + steps, serial_reference=True)
    ts = vkdriver.frame_seconds(warm + steps)[warm:]
    t = float(np.mean(ts))
    modes["serial"] = {"s_per_frame": t, "mtri_s": tris / t / 1e6, "threads": 1, "frames": len(ts)}
    # the same with the shader stage interpreted (round 1's baseline), one frame, for the record
    reference_native_shaders(False)
    vkdriver.run(vkdriver.ICD_REF, scene, frames=1, serial_reference=True)
    ti = vkdriver.frame_seconds(1)[0]
    reference_native_shaders(True)
    interpreted = {"s_per_frame": ti, "mtri_s": tris / ti / 1e6, "threads": 1, "frames": 1}
    best = max(modes, key=lambda k: modes[k]["mtri_s"])
    # headline = the faster mode ("all the host threads it can use"); on the mesh scenes that is the serial one
    head = best
    v = modes[head]["mtri_s"]
    line = {
        "impl": "reference", "metric": "triangle throughput", "value": v, "unit": "Mtri/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": modes[head]["s_per_frame"] * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": workload_config(workload, scene, args.gpus,
                                  PARALLELISM_DESC["single"] if args.gpus <= 1 else
                                  PARALLELISM_DESC["multicast"].format(n=args.gpus), RASTER_PATH_DESC["auto"]),
        "reference_arm": {
            "timing": "host wall clock around vkQueueSubmit + vkQueueWaitIdle of the reference ICD "
                      "(cmd_exec.cpp:187-201 replaying the frame's command buffer)",
            "shader_stage": "native: the reference's LLVM-6 JIT (spirv_compile.cpp) cannot be built here; its "
                            "spirv_compile.h interface is served by natively compiled C++ equivalents of this "
                            "scene's shaders (oracle/ref/ref_glue.cpp, bit-identical to the SPIR-V interpreter "
                            "oracle/spirv_cpu.cpp that serves every other module)",
            "modes": modes, "fastest_mode": best, "serial_with_interpreted_shaders": interpreted},
        "cpu_baseline": {"value": v, "unit": "Mtri/s", "cores": modes[head]["threads"], "kind": "reference",
                         "sample": f"{modes[head]['frames']} full frames through the reference ICD, {head} mode; "
                                   f"serial (1 thread): {modes['serial']['frames']} frames, threaded as shipped "
                                   "(8 threads): 1 frame, see reference_arm",
                         "serial_mtri_s": modes["serial"]["mtri_s"],
                         "threaded_mtri_s": modes.get("threaded", {}).get("mtri_s"), "host_cpus": os.cpu_count()},
        "e2e": {"value": v, "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_frames(gpu, L, submit, steps: int, warmup: int = 3) -> float:
    """mean device time (ms) of `steps` frames on this GPU: L2 flushed before each, CUDA events on the
    library stream around each frame"""
    ms = C.c_float()
    for _ in range(warmup):
        submit()
    gpu.flush()
    times = []
    for _ in range(steps):
        gpu.check(L.vb200_l2_flush(), "l2_flush")
        gpu.check(L.vb200_event_record(2), "event")
        submit()
        gpu.check(L.vb200_event_record(3), "event")
        gpu.check(L.vb200_event_elapsed_ms(2, 3, C.byref(ms)), "elapsed")
        times.append(ms.value)
    return float(np.mean(times))


# ------------------------------------------------------------------------------------------------
def run_ours(args, workload: str) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    if multi:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    gpu = abi.backend("vb200", local)
    L = gpu.lib
    L.vb200_event_elapsed_ms.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.vb200_get_phase_times.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.vb200_set_option.argtypes = [C.c_char_p, C.c_int64]
    L.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mem_upload.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mem_download.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mem_device_ptr.argtypes = [C.c_void_p]
    L.vb200_mem_device_ptr.restype = C.c_void_p
    L.vb200_stream.restype = C.c_void_p
    L.vb200_tiles_pack.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64]
    L.vb200_tiles_unpack.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64, C.c_int]
    L.vb200_tiles_per_rank.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.vb200_tiles_per_rank.restype = C.c_uint32

    if args.raster_path == "ordered":
        gpu.check(L.vb200_set_option(b"raster_path", 1), "set_option")
    scene = build_scene(workload)
    tris = scene.triangles()
    npx = scene.width * scene.height
    fused = multicast = False
    if multi:
        ext = torch.cuda.ExternalStream(L.vb200_stream())
        L.vb200_mgpu_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.vb200_mgpu_info.argtypes = [C.POINTER(C.c_int)] * 3
        L.vb200_mgpu_alloc.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
        L.vb200_mgpu_upload.argtypes = [C.c_void_p, C.c_uint64]
        if args.exchange in ("fused", "fused-p2p"):
            # the product's own multi-GPU layer: bootstrap, symmetric colour image (NVSwitch multicast mapping
            # where available), device-side barrier, sliced input upload. torch.distributed above serves the
            # harness only (timing all-reduce, the NCCL all-gather baseline).
            if args.exchange == "fused-p2p":
                os.environ["VB200_MGPU_NO_MULTICAST"] = "1"
            session = "bench" + os.environ.get("MASTER_PORT", "0")
            gpu.check(L.vb200_mgpu_init(rank, world, local, session.encode()), "mgpu_init")
            mc = C.c_int()
            L.vb200_mgpu_info(None, None, C.byref(mc))
            fused, multicast = True, bool(mc.value)
            gpu.check(L.vb200_set_option(b"mgpu_mirrors", 1), "set_option")    # inputs: symmetric mirrors
            cptr, dptr = C.c_void_p(), C.c_void_p()
            gpu.check(L.vb200_mgpu_alloc(npx * 4, C.byref(cptr)), "mgpu_alloc")
            color_ptr = cptr.value
        else:
            gpu.check(L.vb200_set_tile_owner(rank, world), "set_tile_owner")
            color_t = torch.empty(npx, dtype=torch.int32, device="cuda")
            color_ptr = color_t.data_ptr()
        depth_t = torch.empty(npx, dtype=torch.float32, device="cuda")
        bound = scenes.BoundScene(gpu, scene, color_device_ptr=color_ptr, depth_device_ptr=depth_t.data_ptr())
        if not fused:
            slots = L.vb200_tiles_per_rank(scene.width, scene.height, world)
            send = torch.empty(slots * 4096, dtype=torch.uint8, device="cuda")
            recv = torch.empty(world * slots * 4096, dtype=torch.uint8, device="cuda")
    else:
        bound = scenes.BoundScene(gpu, scene)
    bufs = scene_host_buffers(bound)
    if multi:
        bufs = [(a, is_in) for a, is_in in bufs if is_in]    # attachments live in device memory
    for a, _ in bufs:
        gpu.check(L.vb200_mem_register(a.ctypes.data, a.nbytes), "mem_register")
    in_bytes = sum(a.nbytes for a, is_in in bufs if is_in)
    out_bytes = sum(a.nbytes for a, is_in in bufs if not is_in)

    def exchange():
        """sort-first assemble on the library stream. fused: the tile kernels already stored every pixel
        into all ranks' images over NVLink, only a cross-rank barrier is left. all-gather: owned colour
        tiles -> NCCL all-gather -> un-tile."""
        if fused:
            gpu.check(L.vb200_mgpu_barrier(), "mgpu_barrier")
            return
        with torch.cuda.stream(ext):
            gpu.check(L.vb200_tiles_pack(C.byref(bound.color_img), send.data_ptr(), send.numel()), "tiles_pack")
            dist.all_gather_into_tensor(recv, send)
            gpu.check(L.vb200_tiles_unpack(C.byref(bound.color_img), recv.data_ptr(), recv.numel(), world),
                      "tiles_unpack")

    def barrier():
        gpu.flush()
        if multi:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, device-timed --------------------------------
    gpu.check(L.vb200_set_sync_mode(1), "set_sync_mode")
    for a, is_in in bufs:
        if is_in:
            gpu.check(L.vb200_mem_upload(a.ctypes.data, a.nbytes), "mem_upload")

    def step_device():
        bound.submit()
        if multi:
            exchange()

    ms = C.c_float()
    times = []
    # clocks/throttle reasons are sampled from the warm-up on, so that the short timed region (a few
    # milliseconds) sits inside a window with enough NVML samples
    with ClockSampler(local) as clocks:
        for _ in range(max(args.warmup, 50)):
            step_device()
        barrier()
        gpu.reset_stats()
        for _ in range(args.steps):
            gpu.check(L.vb200_l2_flush(), "l2_flush")
            if multi:
                dist.barrier()
                if fused:
                    # the host barrier releases the ranks tens of microseconds apart; a device-side barrier in front
                    # of the start event lines the GPUs up, so that the step times the frame and not that skew
                    gpu.check(L.vb200_mgpu_barrier(), "mgpu_barrier")
            gpu.check(L.vb200_event_record(0), "event")
            step_device()
            gpu.check(L.vb200_event_record(1), "event")
            gpu.check(L.vb200_event_elapsed_ms(0, 1, C.byref(ms)), "elapsed")
            times.append(ms.value)
    barrier()
    launches = gpu.stats()["kernel_launches"] / max(1, args.steps)
    t_local = float(np.mean(times))
    if multi:
        tt = torch.tensor([t_local], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step = float(tt.item())
    else:
        t_step = t_local

    # ---------------- diagnostics: per-phase device time, fragment counts ------------------------
    L.vb200_set_option(b"time_kernels", 1)
    for _ in range(3):
        gpu.check(L.vb200_l2_flush(), "l2_flush")
        step_device()
    barrier()
    pm = (C.c_double * 5)()
    pc = (C.c_uint64 * 5)()
    L.vb200_get_phase_times(pm, pc, 5)
    phase = {n: (pm[i] / pc[i] if pc[i] else 0.0) for i, n in enumerate(["clear", "vertex", "setup", "bin", "tiles"])}
    phase["clear"] *= (2 if scene.depth else 1)  # colour + depth clears per frame
    L.vb200_set_option(b"time_kernels", 0)
    L.vb200_set_option(b"count_fragments", 1)
    gpu.reset_stats()
    step_device()
    barrier()
    st = gpu.stats()
    if multi:    # fragment / pair counters are per rank (each rank rasterises only its own tiles)
        cnt = torch.tensor([st["fragments_covered"], st["fragments_shaded"], st["tile_pairs"]], device="cuda")
        dist.all_reduce(cnt)
        st["fragments_covered"], st["fragments_shaded"], st["tile_pairs"] = (int(v) for v in cnt.tolist())
    L.vb200_set_option(b"count_fragments", 0)

    # N > 1: what the exchange really moves, read from the GPU's own NVLink counters (NVML, no profiler): payload
    # bytes this rank sent and received per frame over a run of frames long enough for the counters to tick
    nvlink = None
    if multi:
        n_nv = 200
        barrier()
        c0 = nvlink_kib(local)
        for _ in range(n_nv):
            step_device()
        barrier()
        time.sleep(0.05)
        c1 = nvlink_kib(local)
        if c0 and c1:
            nvlink = {"tx_bytes_per_frame": (c1[0] - c0[0]) * 1024 // n_nv, "rx_bytes_per_frame": (c1[1] - c0[1]) * 1024 // n_nv,
                      "frames": n_nv, "rank": rank, "source": "NVML NVLINK_THROUGHPUT_DATA_TX/RX, all links of this rank's GPU"}
        else:
            # (this pool's boxes answer NVML_ERROR_NOT_SUPPORTED for the link counters.) The traffic of the fused
            # exchange is fixed by construction: a rank stores each pixel of the tiles it owns once towards the
            # switch, which replicates it; it receives every other rank's pixels.
            tiles_x, tiles_y = (scene.width + 31) // 32, (scene.height + 31) // 32
            own_px = sum(min(32, scene.width - 32 * (t % tiles_x)) * min(32, scene.height - 32 * (t // tiles_x))
                         for t in range(rank, tiles_x * tiles_y, world)) if fused else 0
            nvlink = {"tx_bytes_per_frame": own_px * 4 if multicast else own_px * 4 * (world - 1),
                      "rx_bytes_per_frame": (npx - own_px) * 4, "rank": rank,
                      "source": "computed from the tile ownership (NVML link counters not supported on this box)"}

    # ---------------- e2e: host buffers through the C-ABI, copies inside the timed region --------
    e2e_steps = max(3, min(args.steps, 10))
    if not multi:
        gpu.check(L.vb200_set_sync_mode(0), "set_sync_mode")
    else:
        # N>1: host<->device traffic is sharded like the frame. Every rank uploads 1/N of each input over
        # ITS PCIe link and the slices are replicated into all ranks' HBM mirrors over NVLink (the product's
        # vb200_mgpu_upload; NCCL all-gather in the baseline mode); after the exchange every rank holds the whole
        # image and copies its band of rows into one host buffer shared by all ranks (POSIX shared memory,
        # page-locked in every process).
        from multiprocessing import shared_memory

        inputs = [a for a, is_in in bufs if is_in]
        h2d_job = sum(a.nbytes for a in inputs)
        shm_name = f"vb200_bench_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            try:
                shm = shared_memory.SharedMemory(name=shm_name, create=True, size=npx * 4)
            except FileExistsError:    # left behind by a run that was killed: take it over
                stale = shared_memory.SharedMemory(name=shm_name)
                stale.close()
                stale.unlink()
                shm = shared_memory.SharedMemory(name=shm_name, create=True, size=npx * 4)
        dist.barrier()
        if rank != 0:
            shm = shared_memory.SharedMemory(name=shm_name)
            try:    # rank 0 owns the segment; keep this process's resource tracker from unlinking it too
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        frame_host = np.ndarray((npx,), dtype=np.int32, buffer=shm.buf)
        if fused:
            gpu.check(L.vb200_set_option(b"mgpu_mirrors", 0), "set_option")    # page-locking only, no mirror needed
        gpu.check(L.vb200_mem_register(frame_host.ctypes.data, frame_host.nbytes), "mem_register(shared frame)")
        lo, hi = tiles.row_band(scene.height, rank, world)
        band = slice(lo * scene.width, hi * scene.width)
        band_bytes = (hi - lo) * scene.width * 4
        if fused:
            # this rank's band of the finished image, as an image of its own for vb200_present
            band_img = abi.Image.from_buffer_copy(bound.color_img)
            band_img.pixels = color_ptr + lo * scene.width * 4
            band_img.height = hi - lo
            L.vb200_present.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        else:
            class _DevBytes:    # raw device range -> torch tensor (no copy)
                def __init__(self, ptr, nbytes):
                    self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                                     "version": 2}

            shards = []    # (host array, slice bytes, gathered device tensor, own slice view)
            for a in inputs:
                sl, _tail = tiles.upload_shard(a.nbytes, world)
                if sl == 0:
                    shards.append((a, 0, None, None))
                    continue
                dev = L.vb200_mem_device_ptr(a.ctypes.data)
                full = torch.as_tensor(_DevBytes(dev, sl * world), device="cuda")
                shards.append((a, sl, full, full[rank * sl:(rank + 1) * sl]))
            band_host = torch.from_numpy(frame_host[band])

    def step_e2e():
        """one frame as an application sees it: inputs come from host memory, the finished frame ends
        up in host memory."""
        if multi and fused:
            # every rank sends 1/N of each input over its own PCIe link and replicates the slice to the other
            # ranks over NVLink (vb200_mgpu_upload); one barrier later all mirrors are complete — and every
            # rank has finished copying out the previous frame, so the tile kernels may store into peers again
            for a in inputs:
                gpu.check(L.vb200_mgpu_upload(a.ctypes.data, a.nbytes), "mgpu_upload")
            gpu.check(L.vb200_mgpu_barrier(), "mgpu_barrier")
            bound.submit()
            exchange()
            t = C.c_int()
            gpu.check(L.vb200_present(C.byref(band_img), frame_host.ctypes.data + lo * scene.width * 4, band_bytes,
                                      C.byref(t)), "present")
            gpu.flush()
            return
        if multi:
            for a, sl, full, mine in shards:
                if sl == 0:
                    gpu.check(L.vb200_mem_upload(a.ctypes.data, a.nbytes), "mem_upload")
                    continue
                gpu.check(L.vb200_mem_upload(a.ctypes.data + rank * sl, sl), "mem_upload")
                if a.nbytes > sl * world:    # remainder of the division: a few hundred bytes, every rank
                    gpu.check(L.vb200_mem_upload(a.ctypes.data + sl * world, a.nbytes - sl * world), "mem_upload")
                with torch.cuda.stream(ext):
                    dist.all_gather_into_tensor(full, mine)
        bound.submit()
        if multi:
            exchange()
            with torch.cuda.stream(ext):
                band_host.copy_(color_t[band], non_blocking=True)
        gpu.flush()

    for _ in range(2):
        step_e2e()
    barrier()
    gpu.reset_stats()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    if multi:
        torch.cuda.synchronize()
    t_e2e_local = (time.perf_counter() - t0) / e2e_steps
    st2 = gpu.stats()
    if multi:    # whole-job bytes (all ranks together move each input and the image exactly once)
        st2["h2d_bytes"] = h2d_job * e2e_steps
        st2["d2h_bytes"] = npx * 4 * e2e_steps
    if multi:
        tt = torch.tensor([t_e2e_local], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    else:
        t_e2e = t_e2e_local
    single = None
    if multi:
        barrier()    # nobody tears down memory that peers may still be storing into
        final = frame_host.copy().view(np.uint8).reshape(scene.height, scene.width, 4)
        img_hash = scenes.image_hash(final, None)
        # the same workload on ONE of these GPUs (rank 0, tile ownership and exchange off, inputs resident):
        # the denominator of the parallel efficiency, measured in the same run on the same box
        if rank == 0:
            gpu.check(L.vb200_set_tile_owner(0, 1), "set_tile_owner")
            t1 = time_frames(gpu, L, bound.submit, max(3, min(args.steps, 10)))
            single = {"value": tris / (t1 * 1e-3) / 1e6, "unit": "Mtri/s", "ms_per_step": t1, "n_gpus": 1,
                      "note": "same workload, rank 0 alone, same run"}
            gpu.flush()
        L.vb200_mem_unregister.argtypes = [C.c_void_p]
        L.vb200_mem_unregister(frame_host.ctypes.data)
        if not fused:
            del band_host
        del frame_host
        dist.barrier()
        shm.close()
        if rank == 0:
            shm.unlink()
        dist.destroy_process_group()
        if rank != 0:
            return
    else:
        img_hash = scenes.image_hash(bound.color, bound.depth)

    # ---------------- e2e with frames in flight (N = 1): the present path ------------------------
    # Same per-frame host traffic as `e2e` (every input uploaded, the finished colour image copied to host
    # memory) but the copy-out runs on the library's second stream (vb200_present) while the next frame
    # is uploaded and rendered into the other of two colour images — what a double-buffered application
    # gets. Depth stays in HBM (a present shows colour). Reported next to the synchronous `e2e`.
    pipelined = None
    if not multi:
        L.vb200_present.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        L.vb200_present_wait.argtypes = [C.c_int]
        L.vb200_mem_unregister.argtypes = [C.c_void_p]
        gpu.check(L.vb200_set_sync_mode(1), "set_sync_mode")
        col2 = np.zeros_like(bound.color)
        gpu.check(L.vb200_mem_register(col2.ctypes.data, col2.nbytes), "mem_register")
        frames = [bound, scenes.BoundScene(gpu, scene, color=col2, depth=bound.depth)]
        shown = [np.zeros_like(bound.color) for _ in range(2)]
        for a in shown:
            gpu.check(L.vb200_mem_register(a.ctypes.data, a.nbytes), "mem_register")
        inputs = [a for a, is_in in bufs if is_in]
        tickets = [0, 0]

        def frame(i):
            j = i & 1
            if tickets[j]:
                gpu.check(L.vb200_present_wait(tickets[j]), "present_wait")    # host image j is free again
            for a in inputs:
                gpu.check(L.vb200_mem_upload(a.ctypes.data, a.nbytes), "mem_upload")
            frames[j].submit()
            t = C.c_int()
            gpu.check(L.vb200_present(C.byref(frames[j].color_img), shown[j].ctypes.data, shown[j].nbytes,
                                      C.byref(t)), "present")
            tickets[j] = t.value

        for i in range(4):
            frame(i)
        gpu.flush()
        tickets = [0, 0]
        n_pipe = max(6, 2 * e2e_steps)
        t0 = time.perf_counter()
        for i in range(n_pipe):
            frame(i)
        gpu.flush()
        t_pipe = (time.perf_counter() - t0) / n_pipe
        ok = all(np.array_equal(a, bound.color) for a in shown)
        pipelined = {"value": tris / t_pipe / 1e6, "unit": "Mtri/s", "ms_per_step": t_pipe * 1e3,
                     "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(bound.color.nbytes),
                     "frames_in_flight": 2, "steps": n_pipe,
                     "note": "inputs uploaded every frame; colour image copied out by vb200_present on a second "
                             "stream while the next frame runs; depth stays in HBM",
                     "images_match_synchronous_frame": bool(ok)}
        for a in shown + [col2]:
            L.vb200_mem_unregister(a.ctypes.data)
        gpu.check(L.vb200_set_sync_mode(0), "set_sync_mode")

    # N=1 run of the default workload: the scaling runs (N>1) use the 8K config, so its single-GPU time
    # is reported next to the headline for a same-workload scaling curve
    scaling_base = None
    if not multi and getattr(args, "auto_workload", False) and workload != "c5":
        gpu.check(L.vb200_set_sync_mode(1), "set_sync_mode")    # inputs resident in HBM, as for `value`
        sc5 = build_scene("c5")
        b5 = scenes.BoundScene(gpu, sc5)
        for a, is_in in scene_host_buffers(b5):
            gpu.check(L.vb200_mem_register(a.ctypes.data, a.nbytes), "mem_register")
            if is_in:
                gpu.check(L.vb200_mem_upload(a.ctypes.data, a.nbytes), "mem_upload")
        t5 = time_frames(gpu, L, b5.submit, max(3, min(args.steps, 10)))
        gpu.flush()
        scaling_base = {"workload": WORKLOAD_DESC["c5"], "value": sc5.triangles() / (t5 * 1e-3) / 1e6,
                        "unit": "Mtri/s", "ms_per_step": t5, "n_gpus": 1,
                        "note": "the N>1 runs of this bench use this workload (sort-first over screen tiles)"}
        L.vb200_mem_unregister.argtypes = [C.c_void_p]
        for a, _ in scene_host_buffers(b5):
            L.vb200_mem_unregister(a.ctypes.data)
        del b5, sc5

    # ---------------- CPU baseline (rank 0, N = 1): the reference's own rasterizer ----------------
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        # N>1: serial mode only (a threaded 8K frame takes ~25 s); the colour image is what is compared
        r = time_reference(scene, 2 if not multi else 1, 0, threaded_steps=0 if multi else 1,
                           hash_depth=not multi, interpreted_frames=0 if multi else 1)
        if r is not None:
            best = max((k for k in ("serial", "threaded") if k in r), key=lambda k: r[k]["mtri_s"])
            cpu = {"value": r[best]["mtri_s"], "unit": "Mtri/s", "cores": r[best]["threads"], "kind": "reference",
                   "sample": f"2 full frames of the same scene, best mode = {best}; serial "
                             f"{r['serial']['mtri_s']:.3f} Mtri/s (1 thread), threaded "
                             f"{r.get('threaded', {}).get('mtri_s', float('nan')):.3f} Mtri/s (8 threads as shipped)",
                   "shader_stage": "native C++ equivalents of the scene's shaders (oracle/ref/ref_glue.cpp); with the "
                                   "SPIR-V interpreter instead: "
                                   f"{r.get('serial_interpreted', {}).get('mtri_s', float('nan')):.3f} Mtri/s",
                   "host_cpus": os.cpu_count()}
            parity = "bit-exact vs reference serial path" if r["hash"] == img_hash else "MISMATCH vs reference"

    L.vb200_last_tile_kernel.restype = C.c_char_p
    tile_kernel = (L.vb200_last_tile_kernel() or b"").decode()
    peak, peak_src = measured_peak_gbs()
    px = scene.width * scene.height
    per_px = 8 if scene.depth else 4
    tile_bytes = px * per_px
    tiles_ms = phase["tiles"]
    achieved = tile_bytes / (tiles_ms * 1e-3) / 1e9 if tiles_ms > 0 else 0.0
    frame_bytes = scene.algorithmic_bytes()
    frame_gbs = frame_bytes / (t_step * 1e-3) / 1e9
    # secondary limiter: issue-slot utilisation of the tile kernel = warp instructions per
    # launch (ncu) / its live duration / (4 schedulers x SMs x SM clock)
    issue_info = None
    winstr = NCU_WARP_INSTRUCTIONS.get((workload, world, tile_kernel))
    if winstr and tiles_ms > 0:
        sm_mhz = clocks.summary().get("sm_mhz") or 1965.0
        peak_issue = 4 * 148 * sm_mhz * 1e6
        issue_info = {"warp_instructions": winstr, "achieved_ginstr_s": winstr / (tiles_ms * 1e-3) / 1e9,
                      "peak_ginstr_s": peak_issue / 1e9, "frac": winstr / (tiles_ms * 1e-3) / peak_issue,
                      "source": NCU_TRAFFIC_SOURCE}
    line = {
        "metric": "triangle throughput", "value": tris / (t_step * 1e-3) / 1e6, "unit": "Mtri/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step,
        "higher_is_better": True, "scaling": "strong" if multi else "weak", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": workload_config(
            workload, scene, world,
            (PARALLELISM_DESC["multicast" if multicast else "p2p" if fused else "allgather"].format(n=world)
             if multi else PARALLELISM_DESC["single"]),
            RASTER_PATH_DESC["ordered" if args.raster_path == "ordered" else "auto"]),
        "gfrag_s": st["fragments_covered"] / (t_step * 1e-3) / 1e9,
        "fragments": {"covered": st["fragments_covered"], "shaded": st["fragments_shaded"],
                      "triangles_out": st["triangles_out"], "tile_pairs": st["tile_pairs"]},
        "phase_ms": phase,
        # N > 1: `achieved` is the whole job's (all ranks' tile kernels together write the frame once), so the
        # fraction is taken against the N GPUs' combined peak
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "n_gpus": world,
                     "frac": achieved / (peak * world), "traffic": NCU_TRAFFIC.get((workload, world, tile_kernel)),
                     "traffic_source": NCU_TRAFFIC_SOURCE if (workload, world, tile_kernel) in NCU_TRAFFIC else None,
                     "kernel": tile_kernel,
                     "algorithmic_bytes": tile_bytes, "kernel_ms": tiles_ms, "peak_source": peak_src,
                     "frame": {"algorithmic_bytes": frame_bytes, "achieved": frame_gbs,
                               "frac": frame_gbs / (peak * world)},
                     "issue": issue_info,
                     # (not HBM: the SM resources the kernel saturates first, from the committed ncu captures)
                     "sm_limiters_ncu": NCU_SM_LIMITERS.get((workload, world, tile_kernel))},
        "e2e": {"value": tris / t_e2e / 1e6, "unit": "Mtri/s", "ms_per_step": t_e2e * 1e3,
                "h2d_bytes_per_step": st2["h2d_bytes"] // e2e_steps, "d2h_bytes_per_step": st2["d2h_bytes"] // e2e_steps,
                "steps": e2e_steps, "host_buffers": {"inputs": in_bytes, "attachments": out_bytes}},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    if len(scene.draws) > 1 and not multi:
        # many-draw frames: the timed step contains the host-side recording of every draw (this harness calls the
        # C-ABI from Python through ctypes); split it into the time the host spends issuing the calls, what an
        # empty ctypes call costs, and the device time of the batch (sum of the kernel phases)
        n_draws = len(scene.draws)
        t0 = time.perf_counter()
        bound.submit()
        t_rec = time.perf_counter() - t0
        gpu.check(L.vb200_flush(), "flush")
        t0 = time.perf_counter()
        for _ in range(n_draws):
            L.vb200_abi_version()
        t_null = time.perf_counter() - t0
        line["many_draws"] = {"draws": n_draws, "host_record_ms": t_rec * 1e3, "host_record_us_per_draw": t_rec / n_draws * 1e6,
                              "empty_ctypes_call_us": t_null / n_draws * 1e6,
                              "device_ms": sum(v for k, v in phase.items()),
                              "note": "ms_per_step = host recording of all draws + one batch on the device"}
        if vkdriver.available() and os.path.exists(vkdriver.ICD_CUDA):
            # the same command buffer through the CUDA ICD (visor's own vkCmdDrawIndexed recording, C++ replay, no
            # Python per draw; coherent memory: uploads and read-back inside), next to the mesh as ONE draw
            gpu.check(L.vb200_set_sync_mode(0), "set_sync_mode")
            import dataclasses
            one = dataclasses.replace(scene, draws=[dataclasses.replace(scene.draws[0], first=0, count=sum(d.count for d in scene.draws))])
            icd = {}
            for tag, sc_ in (("many_draws", scene), ("one_draw", one)):
                vkdriver.run(vkdriver.ICD_CUDA, sc_, frames=6)
                icd[tag + "_ms"] = float(np.mean(vkdriver.frame_seconds(6)[1:])) * 1e3
            line["many_draws"]["cuda_icd_queue_submit"] = icd
    if pipelined:
        line["e2e_pipelined"] = pipelined
    if nvlink:
        line["nvlink"] = nvlink
    if single:
        line["single_gpu_same_workload"] = single
    if scaling_base:
        line["scaling_base"] = scaling_base
    if cpu:
        line["cpu_baseline"] = cpu
    if parity:
        line["parity"] = parity
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c1", "c2", "c3", "c4", "c5", "c6"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="fused", choices=["fused", "fused-p2p", "allgather"],
                    help="N>1: fused = tile kernels store into every rank's image over NVLink (one NVSwitch "
                         "multicast store per pixel when an NVLS mapping is available, else one store per peer); "
                         "fused-p2p = force per-peer stores; allgather = pack + NCCL all-gather + unpack")
    ap.add_argument("--raster-path", default="auto", choices=["auto", "ordered"],
                    help="diagnostic: force the in-order tile kernel even for order-independent passes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    workload = args.workload
    args.auto_workload = workload == "auto"
    if workload == "auto":
        workload = "c3" if args.gpus <= 1 else "c5"
    if args.impl == "reference":
        run_reference_arm(args, workload)
    else:
        run_ours(args, workload)


if __name__ == "__main__":
    main()
