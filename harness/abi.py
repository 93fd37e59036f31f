"""ctypes mirror of include/visor_b200.h and a Backend wrapper.

The same struct layouts drive three shared libraries that export the same operator set under
different prefixes:
  vb200_  visor_b200/libvisor_b200.so   the CUDA product
  vor_    oracle/libvisor_oracle.so     our CPU restatement (test oracle)
  vref_   oracle/_ref/libvisor_ref.so   the reference's own rasterizer/texture unit, compiled unmodified
Tests feed identical inputs to each and compare bytes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# VkFormat / state enum values (3rdparty/vulkan.h, header 42)
FMT_R8_UNORM = 9
FMT_R8G8B8A8_UNORM = 37
FMT_B8G8R8A8_UNORM = 44
FMT_R32_SFLOAT = 100
FMT_R32_SINT = 99
FMT_R32G32_SFLOAT = 103
FMT_R32G32B32_SFLOAT = 106
FMT_R32G32B32A32_SFLOAT = 109
FMT_D32_SFLOAT = 126
FMT_BC2_UNORM_BLOCK = 135
FMT_BC3_UNORM_BLOCK = 137
TOPO_LIST, TOPO_STRIP = 3, 4
FRONT_CCW, FRONT_CW = 0, 1
CULL_NONE, CULL_FRONT, CULL_BACK = 0, 1, 2
CMP_NEVER, CMP_LESS, CMP_EQUAL, CMP_LEQUAL, CMP_GREATER, CMP_NOTEQUAL, CMP_GEQUAL, CMP_ALWAYS = range(8)
BF_ZERO, BF_ONE, BF_SRC_ALPHA, BF_ONE_MINUS_SRC_ALPHA = 0, 1, 6, 7
BLEND_ADD = 0
INDEX_U16, INDEX_U32 = 0, 1
DESC_COMBINED_IMAGE_SAMPLER, DESC_UNIFORM_BUFFER = 1, 6


class Image(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32),
                ("depth", C.c_uint32), ("image_type", C.c_uint32), ("format", C.c_uint32),
                ("array_layers", C.c_uint32), ("mip_levels", C.c_uint32),
                ("bytes_per_pixel", C.c_uint32)]


class Buffer(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("size", C.c_uint64)]


class VertexAttr(C.Structure):
    _fields_ = [("format", C.c_uint32), ("stride", C.c_uint32), ("offset", C.c_uint32),
                ("vb", C.c_uint32)]


class Pipeline(C.Structure):
    _fields_ = [("vattrs", VertexAttr * 16), ("topology", C.c_uint32), ("front_face", C.c_uint32),
                ("cull_mode", C.c_uint32), ("depth_compare_op", C.c_uint32),
                ("depth_write_enable", C.c_uint32), ("blend_enable", C.c_uint32),
                ("src_color_blend_factor", C.c_uint32), ("dst_color_blend_factor", C.c_uint32),
                ("color_blend_op", C.c_uint32), ("vs", C.c_void_p), ("fs", C.c_void_p)]


class Binding(C.Structure):
    _fields_ = [("set", C.c_uint32), ("binding", C.c_uint32), ("type", C.c_uint32),
                ("is_image", C.c_uint32), ("buffer", Buffer), ("offset", C.c_uint64),
                ("image", Image)]


class _IB(C.Structure):
    _fields_ = [("buffer", Buffer), ("offset", C.c_uint64), ("index_type", C.c_uint32),
                ("_pad", C.c_uint32)]


class _VB(C.Structure):
    _fields_ = [("buffer", Buffer), ("offset", C.c_uint64)]


class DrawState(C.Structure):
    _fields_ = [("ib", _IB), ("vbs", _VB * 4), ("color", Image), ("depth", Image),
                ("pipeline", C.POINTER(Pipeline)), ("bindings", C.POINTER(Binding)),
                ("num_bindings", C.c_uint32), ("_pad", C.c_uint32), ("pushconsts", C.c_uint8 * 128)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "draws", "triangles_in", "triangles_out", "tile_pairs", "fragments_covered",
        "fragments_shaded", "kernel_launches", "h2d_bytes", "d2h_bytes")]


LIB_PATHS = {
    "vb200": os.path.join(ROOT, "visor_b200", "libvisor_b200.so"),
    "vor": os.path.join(ROOT, "oracle", "libvisor_oracle.so"),
    "vref": os.path.join(ROOT, "oracle", "_ref", "libvisor_ref.so"),
}


def np_ptr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


def make_image(arr: Optional[np.ndarray], width: int, height: int, fmt: int, bpp: int = 4,
               layers: int = 1, mips: int = 1) -> Image:
    im = Image()
    im.pixels = np_ptr(arr)
    im.width, im.height, im.depth = width, height, 1
    im.image_type = 1
    im.format = fmt
    im.array_layers, im.mip_levels, im.bytes_per_pixel = layers, mips, bpp
    return im


def make_buffer(arr: Optional[np.ndarray]) -> Buffer:
    b = Buffer()
    b.bytes = np_ptr(arr)
    b.size = 0 if arr is None else arr.nbytes
    return b


class BackendError(RuntimeError):
    pass


class Backend:
    """One of the three libraries behind a uniform Python surface named after the reference's
    operators (gpu.h:59-70, spirv_compile.h:3-10)."""

    def __init__(self, kind: str, init_arg: int = 0) -> None:
        if kind not in LIB_PATHS:
            raise ValueError(kind)
        path = LIB_PATHS[kind]
        if not os.path.exists(path):
            raise BackendError(f"{path} not built (run python -c 'import __graft_entry__ as g; g.build()')")
        self.kind = kind
        self.lib = C.CDLL(path)
        self.p = kind + "_"
        L = self.lib
        self._f("shader_create", C.c_void_p, [C.POINTER(C.c_uint32), C.c_size_t])
        self._f("shader_entry", C.c_void_p, [C.c_void_p, C.c_char_p])
        self._f("shader_destroy", None, [C.c_void_p])
        self._f("clear_color", C.c_int, [C.POINTER(Image), C.POINTER(C.c_float)])
        self._f("clear_depth", C.c_int, [C.POINTER(Image), C.c_float])
        self._f("draw", C.c_int, [C.POINTER(DrawState), C.c_int, C.c_uint32, C.c_int])
        self._f("sample", C.c_int, [C.POINTER(Image), C.c_int, C.c_uint64, C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.c_size_t])
        self._f("flush", C.c_int, [])
        self._f("init", C.c_int, [C.c_int])
        self._f("last_error", C.c_char_p, [])
        if kind != "vref":
            self._f("get_stats", C.c_int, [C.POINTER(Stats)])
            self._f("reset_stats", None, [])
        rc = self.fn("init")(init_arg)
        if rc != 0:
            raise BackendError(f"{kind}_init failed: {self.last_error()}")
        self._keep: List[object] = []

    def _f(self, name: str, restype, argtypes) -> None:
        fn = getattr(self.lib, self.p + name)
        fn.restype = restype
        fn.argtypes = argtypes

    def fn(self, name: str):
        return getattr(self.lib, self.p + name)

    def last_error(self) -> str:
        e = self.fn("last_error")()
        return e.decode() if e else ""

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise BackendError(f"{self.kind}: {what} failed ({rc}): {self.last_error()}")

    # ---- spirv_compile.h
    def CompileFunction(self, words: np.ndarray) -> int:
        w = np.ascontiguousarray(words, dtype=np.uint32)
        h = self.fn("shader_create")(w.ctypes.data_as(C.POINTER(C.c_uint32)), w.size)
        if not h:
            raise BackendError(f"{self.kind}: CompileFunction failed: {self.last_error()}")
        return h

    def GetFuncPointer(self, module: int, name: str = "main") -> int:
        e = self.fn("shader_entry")(module, name.encode())
        if not e:
            raise BackendError(f"{self.kind}: GetFuncPointer({name}) failed: {self.last_error()}")
        return e

    def DestroyFunction(self, module: int) -> None:
        self.fn("shader_destroy")(module)

    # ---- gpu.h
    def ClearTarget(self, image: Image, value) -> None:
        if isinstance(value, (tuple, list, np.ndarray)):
            v = (C.c_float * 4)(*[float(x) for x in value])
            self.check(self.fn("clear_color")(C.byref(image), v), "ClearTarget(colour)")
        else:
            self.check(self.fn("clear_depth")(C.byref(image), float(value)), "ClearTarget(depth)")

    def DrawTriangles(self, state: DrawState, num_verts: int, first: int, indexed: bool) -> None:
        self.check(self.fn("draw")(C.byref(state), int(num_verts), int(first), int(bool(indexed))),
                   "DrawTriangles")

    def flush(self) -> None:
        self.check(self.fn("flush")(), "flush")

    def sample(self, image: Image, uvw: np.ndarray, cube: bool = False, byte_offset: int = 0) -> np.ndarray:
        uvw = np.ascontiguousarray(uvw, dtype=np.float32)
        n = uvw.shape[0]
        out = np.empty((n, 4), dtype=np.float32)
        self.check(self.fn("sample")(C.byref(image), int(cube), byte_offset,
                                     uvw.ctypes.data_as(C.POINTER(C.c_float)),
                                     out.ctypes.data_as(C.POINTER(C.c_float)), n), "sample")
        return out

    def stats(self) -> Dict[str, int]:
        s = Stats()
        self.check(self.fn("get_stats")(C.byref(s)), "get_stats")
        return {n: int(getattr(s, n)) for n, _ in Stats._fields_}

    def reset_stats(self) -> None:
        self.fn("reset_stats")()


_backends: Dict[str, Backend] = {}


def backend(kind: str, init_arg: int = 0) -> Backend:
    """Process-wide singleton per library (the reference's rasterizer state is global)."""
    if kind not in _backends:
        _backends[kind] = Backend(kind, init_arg)
    return _backends[kind]


def available(kind: str) -> bool:
    return os.path.exists(LIB_PATHS[kind])
