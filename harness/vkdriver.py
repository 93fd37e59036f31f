"""ctypes front end of tests/icd/vk_driver.cpp: runs a harness Scene through a real Vulkan call
sequence against an ICD shared library (the reference ICD or the CUDA ICD)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from . import abi
from .scenes import Scene

ROOT = abi.ROOT
DRIVER = os.path.join(ROOT, "oracle", "_ref", "libvk_driver.so")
ICD_REF = os.path.join(ROOT, "oracle", "_ref", "libvisor_ref.so")
ICD_CUDA = os.path.join(ROOT, "integration", "libvisor_b200_icd.so")


class Attr(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("location", "format", "stride", "offset", "binding")]


class Ubo(C.Structure):
    _fields_ = [("set", C.c_uint32), ("binding", C.c_uint32), ("data", C.c_void_p), ("size", C.c_uint64),
                ("offset", C.c_uint64)]


class Tex(C.Structure):
    _fields_ = [("set", C.c_uint32), ("binding", C.c_uint32), ("data", C.c_void_p), ("width", C.c_uint32),
                ("height", C.c_uint32), ("format", C.c_uint32), ("bpp", C.c_uint32), ("layers", C.c_uint32)]


class DrawDesc(C.Structure):
    _fields_ = [("vs_code", C.c_void_p), ("vs_words", C.c_uint32), ("fs_code", C.c_void_p), ("fs_words", C.c_uint32),
                ("num_attrs", C.c_uint32), ("attrs", Attr * 16),
                ("topology", C.c_uint32), ("front_face", C.c_uint32), ("cull_mode", C.c_uint32),
                ("depth_op", C.c_uint32), ("depth_write", C.c_uint32), ("blend_enable", C.c_uint32),
                ("src_factor", C.c_uint32), ("dst_factor", C.c_uint32), ("blend_op", C.c_uint32),
                ("vb", C.c_void_p * 4), ("vb_size", C.c_uint64 * 4), ("vb_offset", C.c_uint64 * 4),
                ("ib", C.c_void_p), ("ib_size", C.c_uint64), ("ib_offset", C.c_uint64), ("index_type", C.c_uint32),
                ("num_ubos", C.c_uint32), ("ubos", Ubo * 4), ("num_tex", C.c_uint32), ("tex", Tex * 4),
                ("push", C.c_uint8 * 128), ("push_size", C.c_uint32),
                ("count", C.c_uint32), ("first", C.c_uint32), ("indexed", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("has_depth", C.c_uint32),
                ("clear_color_enable", C.c_uint32), ("clear_color", C.c_float * 4),
                ("clear_depth_enable", C.c_uint32), ("clear_depth", C.c_float),
                ("num_draws", C.c_uint32), ("draws", C.POINTER(DrawDesc)),
                ("color_out", C.c_void_p), ("depth_out", C.c_void_p)]


def available() -> bool:
    return os.path.exists(DRIVER)


_lib = None


def _driver():
    global _lib
    if _lib is None:
        _lib = C.CDLL(DRIVER)
        _lib.vkd_run.argtypes = [C.c_char_p, C.POINTER(SceneDesc), C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.vkd_run.restype = C.c_int
        _lib.vkd_frame_seconds.argtypes = [C.POINTER(C.c_double), C.c_int]
        _lib.vkd_frame_seconds.restype = C.c_int
    return _lib


def frame_seconds(n: int):
    """host wall clock around vkQueueSubmit + vkQueueWaitIdle of each of the (up to n) frames of the last run()"""
    buf = (C.c_double * max(n, 1))()
    got = _driver().vkd_frame_seconds(buf, n)
    return [buf[i] for i in range(got)]


def run(icd_path: str, scene: Scene, frames: int = 1, serial_reference: bool = True
        ) -> Tuple[np.ndarray, Optional[np.ndarray], float]:
    """Render `scene` through vkCreateInstance .. vkQueueSubmit of the ICD at `icd_path`."""
    w, h = scene.width, scene.height
    color = np.full((h, w, 4), 0xCD, dtype=np.uint8)
    depth = np.full((h, w), 0.75, dtype=np.float32) if scene.depth else None
    keep = []
    draws = (DrawDesc * len(scene.draws))()
    for i, d in enumerate(scene.draws):
        dd = draws[i]
        vs = np.ascontiguousarray(d.pipe.vs, dtype=np.uint32)
        fs = np.ascontiguousarray(d.pipe.fs, dtype=np.uint32)
        keep += [vs, fs]
        dd.vs_code, dd.vs_words, dd.fs_code, dd.fs_words = vs.ctypes.data, vs.size, fs.ctypes.data, fs.size
        dd.num_attrs = len(d.pipe.vattrs)
        for k, (loc, fmt, stride, off, vb) in enumerate(d.pipe.vattrs):
            dd.attrs[k] = Attr(loc, fmt, stride, off, vb)
        dd.topology, dd.front_face, dd.cull_mode = d.pipe.topology, d.pipe.front_face, d.pipe.cull_mode
        dd.depth_op, dd.depth_write = d.pipe.depth_op, int(d.pipe.depth_write)
        if d.pipe.blend is not None:
            dd.blend_enable = 1
            dd.src_factor, dd.dst_factor, dd.blend_op = d.pipe.blend
        for slot, (buf, off) in enumerate(d.vbs):
            dd.vb[slot], dd.vb_size[slot], dd.vb_offset[slot] = buf.ctypes.data, buf.nbytes, off
        if d.ib is not None:
            dd.ib, dd.ib_size, dd.ib_offset, dd.index_type = d.ib[0].ctypes.data, d.ib[0].nbytes, d.ib[1], d.ib[2]
        dd.num_ubos = len(d.ubos)
        for k, (s, b, buf, off) in enumerate(d.ubos):
            dd.ubos[k] = Ubo(s, b, buf.ctypes.data, buf.nbytes, off)
        dd.num_tex = len(d.textures)
        for k, (s, b, data, tw, th, fmt, bpp, layers) in enumerate(d.textures):
            dd.tex[k] = Tex(s, b, data.ctypes.data, tw, th, fmt, bpp, layers)
        if d.push:
            C.memmove(dd.push, d.push, min(len(d.push), 128))
            dd.push_size = min(len(d.push), 128)
        dd.count, dd.first, dd.indexed = d.count, d.first, int(d.indexed)
    sd = SceneDesc()
    sd.width, sd.height, sd.has_depth = w, h, int(scene.depth)
    if scene.clear_color is not None:
        sd.clear_color_enable = 1
        sd.clear_color = (C.c_float * 4)(*scene.clear_color)
    if scene.depth and scene.clear_depth is not None:
        sd.clear_depth_enable = 1
        sd.clear_depth = scene.clear_depth
    sd.num_draws = len(scene.draws)
    sd.draws = C.cast(draws, C.POINTER(DrawDesc))
    sd.color_out = color.ctypes.data
    sd.depth_out = depth.ctypes.data if depth is not None else None
    secs = C.c_double()
    rc = _driver().vkd_run(icd_path.encode(), C.byref(sd), frames, int(serial_reference), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"vk_driver failed with {rc} on {os.path.basename(icd_path)}")
    return color, depth, secs.value
