"""Third opinion on the shader stage: LLVM-generated x86 code for the IR the reference's IRBuilder emits
per opcode (harness/llvm_opinion.py) against the CPU interpreter the oracle uses (oracle/spirv_cpu.cpp),
bit for bit, on the single-op shaders of the known-answer tests. The GPU back end is compared with the
interpreter in tests/test_gpu_parity.py::test_spirv_ops_match_oracle, which closes the triangle."""
import ctypes as C

import numpy as np
import pytest

from harness import abi, shaders

llvmlite = pytest.importorskip("llvmlite")
from harness import llvm_opinion  # noqa: E402

from tests.test_oracle_shader import unit_inputs, unit_state  # noqa: E402

f32 = np.float32
LIBM = ("sin", "cos", "pow")    # LLVM lowers these to libm calls, the interpreter to its own: 1-ulp class


@pytest.fixture(scope="module")
def jit():
    return llvm_opinion.Jit(shaders.UNIT_OPS)


def test_ir_is_unfused_and_has_no_fast_math_flags():
    ir = llvm_opinion.module_ir(shaders.UNIT_OPS)
    for flag in (" fast ", " nnan ", " ninf ", " nsz ", " arcp ", " contract ", " afn ", " reassoc ", "fmuladd", "llvm.fma"):
        assert flag not in ir, flag


@pytest.mark.parametrize("op", shaders.UNIT_OPS)
def test_llvm_codegen_matches_the_interpreter(vor, jit, op):
    verts, ubo = unit_inputs(seed=7, n=96)
    if op in ("sqrt", "invsqrt", "pow"):
        verts[:, 0:4] = np.abs(verts[:, 0:4])
    # special values: signed zeros, infinities, NaN, denormals, ties
    verts[0, 0:4] = [0.0, -0.0, np.inf, -np.inf]
    verts[0, 4:8] = [-0.0, 0.0, 1.0, -1.0]
    verts[1, 0:4] = [np.nan, 1.0, -2.0, 1e-40]
    verts[1, 4:8] = [1.0, np.nan, -2.0, 1e-40]
    verts[2, 4:8] = verts[2, 0:4]
    mod, st, keep = unit_state(vor, op, verts, ubo)
    entry = vor.GetFuncPointer(mod, "main")
    run = vor.lib.vor_run_vertex
    run.argtypes = [C.POINTER(abi.DrawState), C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    out = (C.c_float * 44)()
    for i in range(verts.shape[0]):
        assert run(C.byref(st), entry, i, out) == 0
        got = np.frombuffer(out, dtype=f32)[4:8].copy()
        a, b, c = verts[i, 0:4], verts[i, 4:8], verts[i, 8:12]
        with np.errstate(all="ignore"):
            exp = jit.run(op, a, b, c, ubo[:16], ubo[16:])
        if op in LIBM:
            ok = np.isclose(got, exp, rtol=1e-5, atol=1e-6, equal_nan=True)
            assert ok.all(), (op, i, got, exp)
            continue
        same = got.view(np.uint32) == exp.view(np.uint32)
        # NaN results: the payload/sign of a propagated NaN is not part of either contract
        same |= np.isnan(got) & np.isnan(exp)
        assert same.all(), (op, i, a, b, c, got, exp)
    vor.DestroyFunction(mod)
