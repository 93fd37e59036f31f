#!/usr/bin/env python
"""Developer aid (no GPU needed): JIT-compile every kernel of scaffold.cu for one of the bench shader
pairs with the library's own ptxas and keep the cubins: tools/jit_dump.py c3 /tmp/jit/c3
-> /tmp/jit/c3.<kernel>.cubin (inspect with cuobjdump -sass / -res-usage, nvdisasm -g)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from harness import shaders  # noqa: E402

PAIRS = {
    "c1": (shaders.vs_passthrough, shaders.fs_color),
    "c2": (shaders.vs_mvp_uv, shaders.fs_texture),
    "c3": (lambda: shaders.vs_lit(False), shaders.fs_color),
    "c5": (lambda: shaders.vs_lit(True), shaders.fs_lit_tex),
}


def main():
    name, prefix = sys.argv[1], sys.argv[2]
    os.environ["VB200_DUMP_CUBIN"] = prefix
    os.environ.setdefault("VB200_JIT_VERBOSE", "1")
    import visor_b200
    L = visor_b200.lib()
    L.vb200_shader_create.restype = C.c_void_p
    L.vb200_shader_create.argtypes = [C.c_void_p, C.c_size_t]
    L.vb200_shader_entry.restype = C.c_void_p
    L.vb200_shader_entry.argtypes = [C.c_void_p, C.c_char_p]
    L.vb200_link_check.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.vb200_last_error.restype = C.c_char_p
    ents = []
    for mk in PAIRS[name]:
        w = np.ascontiguousarray(mk(), dtype=np.uint32)
        mod = L.vb200_shader_create(w.ctypes.data, w.size)
        assert mod, L.vb200_last_error()
        ents.append(L.vb200_shader_entry(mod, b"main"))
    sz = C.c_uint64()
    rc = L.vb200_link_check(ents[0], ents[1], C.byref(sz))
    assert rc == 0, L.vb200_last_error()
    print("ok", sz.value)


if __name__ == "__main__":
    main()
