"""The C-ABI library: loads without a GPU, exports every symbol include/visor_b200.h declares, its
struct layouts match the ctypes mirror, shaders lower to unfused IEEE PTX that ptxas accepts for
sm_100a, errors are reported (never thrown), and there is NO CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from harness import abi, shaders

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "visor_b200.h")


@pytest.fixture(scope="module")
def lib():
    import visor_b200
    L = visor_b200.lib()
    L.vb200_shader_create.restype = C.c_void_p
    L.vb200_shader_create.argtypes = [C.c_void_p, C.c_size_t]
    L.vb200_shader_entry.restype = C.c_void_p
    L.vb200_shader_entry.argtypes = [C.c_void_p, C.c_char_p]
    L.vb200_shader_destroy.argtypes = [C.c_void_p]
    L.vb200_entry_ptx.restype = C.c_char_p
    L.vb200_entry_ptx.argtypes = [C.c_void_p]
    L.vb200_entry_stage.argtypes = [C.c_void_p]
    L.vb200_last_error.restype = C.c_char_p
    L.vb200_link_check.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    return L


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"VB200_API\s+[^;(]*?\b(vb200_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in visor_b200.h but not exported: {missing}"
    assert lib.vb200_abi_version() == 1


def test_struct_layouts_match_header():
    """sizeof/offsetof from a C compiler vs the ctypes mirror in harness/abi.py"""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "visor_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(vb200_image), sizeof(vb200_buffer), sizeof(vb200_vertex_attr),
         sizeof(vb200_pipeline), sizeof(vb200_binding), sizeof(vb200_draw_state), sizeof(vb200_stats));
  printf("%zu %zu %zu %zu %zu\n", offsetof(vb200_draw_state, vbs), offsetof(vb200_draw_state, color),
         offsetof(vb200_draw_state, pipeline), offsetof(vb200_draw_state, pushconsts), offsetof(vb200_pipeline, vs));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"),
                        os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[:7] == [C.sizeof(abi.Image), C.sizeof(abi.Buffer), C.sizeof(abi.VertexAttr), C.sizeof(abi.Pipeline),
                         C.sizeof(abi.Binding), C.sizeof(abi.DrawState), C.sizeof(abi.Stats)]
    assert sizes[7:] == [abi.DrawState.vbs.offset, abi.DrawState.color.offset, abi.DrawState.pipeline.offset,
                         abi.DrawState.pushconsts.offset, abi.Pipeline.vs.offset]
    assert C.sizeof(abi.Image) == 40  # VkImage_T is 40 bytes too


def _entry(lib, words):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    mod = lib.vb200_shader_create(w.ctypes.data, w.size)
    assert mod, lib.vb200_last_error()
    e = lib.vb200_shader_entry(mod, b"main")
    assert e
    return mod, e


PAIRS = {
    "c1": (shaders.vs_passthrough, shaders.fs_color),
    "c2": (shaders.vs_mvp_uv, shaders.fs_texture),
    "c3": (lambda: shaders.vs_lit(False), shaders.fs_color),
    "c5": (lambda: shaders.vs_lit(True), shaders.fs_lit_tex),
    "kitchen_sink": (shaders.vs_kitchen_sink, shaders.fs_kitchen_sink),
    "cube": (shaders.vs_pos_dir, shaders.fs_cube),
}


@pytest.mark.parametrize("name", sorted(PAIRS))
def test_shaders_lower_to_ieee_ptx_and_link_for_sm100a(lib, name):
    vs, fs = PAIRS[name]
    m1, ve = _entry(lib, vs())
    m2, fe = _entry(lib, fs())
    assert lib.vb200_entry_stage(ve) == 0 and lib.vb200_entry_stage(fe) == 4
    for e in (ve, fe):
        ptx = lib.vb200_entry_ptx(e).decode()
        assert ".target sm_100a" in ptx
        # arithmetic contract: explicit .rn, never fused, never flushed
        assert not re.search(r"\b(fma|mad)\.(rn\.)?f32", ptx)
        assert ".ftz" not in ptx
        for ins in re.findall(r"^\s*(add|sub|mul|div)\.(\S*)f32", ptx, flags=re.M):
            assert ins[1] == "rn.", ins
    sz = C.c_uint64()
    rc = lib.vb200_link_check(ve, fe, C.byref(sz))
    assert rc == 0, lib.vb200_last_error()
    assert sz.value > 10000
    lib.vb200_shader_destroy(m1)
    lib.vb200_shader_destroy(m2)


def test_shader_and_helpers_are_inlined_into_the_kernels(lib, tmp_path, monkeypatch):
    """ptxas keeps `.func` calls as calls; ptx_inline.cpp inlines the shader entry point and the helpers it calls
    (attribute fetch, texture unit) on the PTX text instead. The cubins of a textured pipeline must not contain a
    call to, or a copy of, any of them — only the IEEE division / reciprocal slow paths remain functions."""
    import shutil
    import subprocess
    if shutil.which("nvdisasm") is None:
        pytest.skip("nvdisasm not on PATH")
    monkeypatch.setenv("VB200_DUMP_CUBIN", str(tmp_path / "k"))
    m1, ve = _entry(lib, shaders.vs_lit(True))
    m2, fe = _entry(lib, shaders.fs_lit_tex())
    sz = C.c_uint64()
    assert lib.vb200_link_check(ve, fe, C.byref(sz)) == 0, lib.vb200_last_error()
    cubins = sorted(tmp_path.glob("k.*.cubin"))
    assert len(cubins) >= 7        # vertex, ordered, five resolve modes
    for c in cubins:
        sass = subprocess.run(["nvdisasm", str(c)], capture_output=True, text=True, check=True).stdout
        funcs = re.findall(r"\.type\s+(\S+),@function", sass)
        extra = [f for f in funcs if not f.startswith("vb200_k_") and "slowpath" not in f]
        assert not extra, (c.name, extra)
        calls = [l for l in sass.splitlines() if re.search(r"\bCALL\b", l) and "slowpath" not in l]
        assert not calls, (c.name, calls[:3])
    # ... and with the switch that keeps the calls (A/B measurements) the functions are back
    monkeypatch.setenv("VB200_JIT_NO_INLINE", "1")
    monkeypatch.setenv("VB200_DUMP_CUBIN", str(tmp_path / "n"))
    m3, ve2 = _entry(lib, shaders.vs_lit(True))
    m4, fe2 = _entry(lib, shaders.fs_lit_tex())
    assert lib.vb200_link_check(ve2, fe2, C.byref(sz)) == 0, lib.vb200_last_error()
    sass = subprocess.run(["nvdisasm", str(tmp_path / "n.vb200_k_tile_ordered.cubin")], capture_output=True, text=True,
                          check=True).stdout
    assert "vb200_fs" in sass and "vb200_sample_tex" in sass
    for m in (m1, m2, m3, m4):
        lib.vb200_shader_destroy(m)


@pytest.mark.parametrize("op", shaders.UNIT_OPS + shaders.MEM_UNIT_OPS)
def test_unit_op_shaders_compile(lib, op):
    mod, e = _entry(lib, shaders.vs_unit(op))
    assert b"vb200_vs" in lib.vb200_entry_ptx(e)
    lib.vb200_shader_destroy(mod)


def test_compile_errors_return_null_with_message(lib):
    """CompileFunction == NULL -> VK_ERROR_DEVICE_LOST in the reference (shaders.cpp:13-14)."""
    good = shaders.vs_passthrough()
    bad_magic = good.copy()
    bad_magic[0] = 0xDEADBEEF
    assert not lib.vb200_shader_create(bad_magic.ctypes.data, bad_magic.size)
    assert b"magic" in lib.vb200_last_error()
    newer = good.copy()
    newer[1] = 0x00010300  # SPIR-V 1.3 > 1.1 (spirv_compile.cpp:652)
    assert not lib.vb200_shader_create(newer.ctypes.data, newer.size)
    # an opcode outside the reference's subset: OpSelect (169)
    from harness.spvasm import Module, Op, VERTEX
    m = Module()
    v4, bl = m.t_fvec(4), m.t_bool()
    a = m.input(v4, 0)
    gl = m.per_vertex_out()
    void = m.t_void()
    f, _ = m.begin_function(void, m.t_func(void))
    m.label()
    x = m.load(v4, a)
    c = m.inst(Op.FOrdLessThan, m.t_vec(bl, 4), x, x)
    m.inst(Op.Select, v4, c, x, x)
    m.ret()
    m.end_function()
    m.entry_point(VERTEX, f, "main", [a, gl])
    w = m.words()
    assert not lib.vb200_shader_create(w.ctypes.data, w.size)
    assert b"Unhandled SPIR-V opcode 169" in lib.vb200_last_error()
    truncated = good[:8].copy()  # cuts OpMemoryModel (3 words) in half
    assert not lib.vb200_shader_create(truncated.ctypes.data, truncated.size)
    mod, _ = _entry(lib, good)
    assert not lib.vb200_shader_entry(mod, b"does_not_exist")
    lib.vb200_shader_destroy(mod)


def test_oracle_and_product_accept_the_same_modules(lib, vor):
    for name, (vs, fs) in PAIRS.items():
        for words in (vs(), fs()):
            mod = vor.CompileFunction(words)
            vor.DestroyFunction(mod)


def test_no_cpu_fallback_without_a_device(lib):
    """On a machine without CUDA the compute entry points must fail loudly, not emulate."""
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is present")
    assert lib.vb200_init(0) == -1  # VB200_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.vb200_last_error()
    img = abi.make_image(np.zeros((4, 4, 4), np.uint8), 4, 4, abi.FMT_B8G8R8A8_UNORM)
    col = (C.c_float * 4)(0, 0, 0, 1)
    assert lib.vb200_clear_color(C.byref(img), col) == -1
    assert lib.vb200_flush() == -1
    # ... and so must everything added next to the draw path: device copies, residency, present
    buf = abi.make_buffer(np.zeros(64, np.uint8))
    lib.vb200_copy_buffer.argtypes = [C.POINTER(abi.Buffer), C.c_uint64, C.POINTER(abi.Buffer), C.c_uint64, C.c_uint64]
    lib.vb200_copy_buffer_to_image.argtypes = [C.POINTER(abi.Buffer), C.c_uint64, C.POINTER(abi.Image), C.c_uint32,
                                               C.c_uint32]
    lib.vb200_present.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    lib.vb200_mem_set_device_local.argtypes = [C.c_void_p, C.c_int]
    lib.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
    ticket = C.c_int()
    host = np.zeros(64, np.uint8)
    assert lib.vb200_copy_buffer(C.byref(buf), 0, C.byref(buf), 32, 16) == -1
    assert lib.vb200_copy_buffer_to_image(C.byref(buf), 0, C.byref(img), 0, 0) == -1
    assert lib.vb200_present(C.byref(img), host.ctypes.data, host.nbytes, C.byref(ticket)) == -1
    assert lib.vb200_present_wait(1) == -1
    assert lib.vb200_mem_register(host.ctypes.data, host.nbytes) == -1
    assert lib.vb200_mem_set_device_local(host.ctypes.data, 1) == -1


def test_byte_over_255_in_two_operations_is_the_ieee_quotient():
    """raster_common.cuh vb200_unorm8: fma(b, r, b*e) with r = RN(1/255), e = RN(1/255 - r) must be the correctly
    rounded float32 quotient b/255 (what the reference's `float(byte) / 255.0f` yields) for every byte value.
    Checked with exact rationals; the GPU side compares all 256 against the IEEE division in
    test_unorm8_conversion_is_the_ieee_quotient."""
    from fractions import Fraction

    def rn32(x):    # Fraction -> nearest float32, ties to even
        c = np.float32(float(x))
        cands = [np.nextafter(c, np.float32(-np.inf)), c, np.nextafter(c, np.float32(np.inf))]
        return min(cands, key=lambda v: (abs(Fraction(float(v)) - x), int(np.float32(v).view(np.uint32)) & 1))

    src = open(os.path.join(os.path.dirname(HEADER), "..", "visor_b200", "csrc", "raster_common.cuh")).read()
    r = re.search(r"const float r = ([0-9.eE+-]+)f;", src).group(1)
    e = re.search(r"const float e = ([0-9.eE+-]+)f;", src).group(1)
    R, E = Fraction(float(np.float32(r))), Fraction(float(np.float32(e)))
    assert np.float32(r) == np.float32(1.0) / np.float32(255.0)
    for b in range(256):
        t = Fraction(float(rn32(b * E)))
        got = rn32(b * R + t)
        want = np.float32(b) / np.float32(255.0)
        assert np.float32(got).view(np.uint32) == want.view(np.uint32), b


def test_tile_kernels_keep_their_occupancy(lib, tmp_path, monkeypatch):
    """The tile kernels are tuned to a number of resident CTAs per SM: the resolve kernels to 8 (64 registers x 128
    threads, <= 27 KB of shared memory each), the ordered kernel to 5 (48 registers x 256 threads). A change that
    silently costs a register or a few KB halves nothing visibly but loses 10-20 % on the GPU; pin it where no GPU
    is needed (C3's and C5's shaders, cuobjdump -res-usage of the JIT cubins)."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    for tag, vs, fs in (("c3", shaders.vs_lit(False), shaders.fs_color()), ("c5", shaders.vs_lit(True), shaders.fs_lit_tex())):
        monkeypatch.setenv("VB200_DUMP_CUBIN", str(tmp_path / tag))
        m1, ve = _entry(lib, vs)
        m2, fe = _entry(lib, fs)
        sz = C.c_uint64()
        assert lib.vb200_link_check(ve, fe, C.byref(sz)) == 0, lib.vb200_last_error()
        for kernel, max_regs, max_smem in (("vb200_k_tile_resolve_min_first", 64, 27 * 1024),
                                           ("vb200_k_tile_resolve_last_wins", 64, 32 * 1024),
                                           ("vb200_k_tile_ordered", 64 if tag == "c5" else 48, 40 * 1024),
                                           ("vb200_k_vertex", 40, 0)):
            out = subprocess.run(["cuobjdump", "-res-usage", str(tmp_path / f"{tag}.{kernel}.cubin")], capture_output=True,
                                 text=True, check=True).stdout
            regs = int(re.search(r"REG:(\d+)", out).group(1))
            smem = int(re.search(r"SHARED:(\d+)", out).group(1))
            assert regs <= max_regs, (tag, kernel, regs)
            assert smem <= max_smem, (tag, kernel, smem)
        lib.vb200_shader_destroy(m1)
        lib.vb200_shader_destroy(m2)
