"""A third, LLVM-generated opinion on the shader stage (test infrastructure).

The reference lowers SPIR-V to LLVM IR with an IRBuilder and JIT-compiles it for x86
(spirv_compile.cpp:645-2432, LLVM 6). That JIT cannot be built here, so the two shader back ends of
this repo — the PTX emitter (visor_b200/csrc/spirv_ptx.cpp) and the CPU interpreter the oracle uses
(oracle/spirv_cpu.cpp) — are both restatements. This module restates the reference a third time at a
different level: for each single-op shader of harness/shaders.py:vs_unit it writes down the *LLVM IR*
the reference's IRBuilder calls produce (same instructions, same vector types, same order; the C
helper functions of spirv_compile.cpp:423-492 as the scalar fmul/fadd chains a compiler without
fast-math emits for them), lets the LLVM in this image (llvmlite) generate x86 code for it and runs
it. What is pinned that way is LLVM's own semantics of those instructions — vector fdiv, fcmp
ordered predicates + select, fsub from -0.0, shufflevector index rules, llvm.sqrt — rather than a
hand-written scalar loop's.

Only the instruction sequence per op is transcribed (cited line by line); module set-up, the
input/output wrappers and the SPIR-V parser are not: the functions here take the operand values
directly.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict

import numpy as np

V4 = "<4 x float>"
V3 = "<3 x float>"
MAT = "[4 x <4 x float>]"


class _Fn:
    """tiny SSA text builder"""

    def __init__(self) -> None:
        self.lines = []
        self.n = 0

    def v(self, text: str) -> str:
        self.n += 1
        name = f"%t{self.n}"
        self.lines.append(f"  {name} = {text}")
        return name

    def do(self, text: str) -> None:
        self.lines.append("  " + text)


def _splat(f: _Fn, s: str, n: int = 4) -> str:
    # IRBuilder::CreateVectorSplat: insertelement into undef at 0, then a zero-mask shufflevector
    ty = f"<{n} x float>"
    ins = f.v(f"insertelement {ty} undef, float {s}, i32 0")
    return f.v(f"shufflevector {ty} {ins}, {ty} undef, <{n} x i32> zeroinitializer")


def _dot(f: _Fn, a: str, b: str, n: int) -> str:
    # CreateDot, spirv_compile.cpp:629-643: ((a0*b0 + a1*b1) + a2*b2) + a3*b3
    ty = f"<{n} x float>"
    acc = f.v(f"fmul float {f.v(f'extractelement {ty} {a}, i64 0')}, {f.v(f'extractelement {ty} {b}, i64 0')}")
    for i in range(1, n):
        p = f.v(f"fmul float {f.v(f'extractelement {ty} {a}, i64 {i}')}, {f.v(f'extractelement {ty} {b}, i64 {i}')}")
        acc = f.v(f"fadd float {acc}, {p}")
    return acc


def _construct4(f: _Fn, elems) -> str:
    # OpCompositeConstruct of scalars, :1774-1788: insertelement chain into undef
    cur = "undef"
    for i, e in enumerate(elems):
        cur = f.v(f"insertelement {V4} {cur}, float {e}, i32 {i}")
    return cur


def _xyz(f: _Fn, a: str) -> str:
    # OpVectorShuffle a, a, 0 1 2 (:1789-1833): equal operand sizes, plain shufflevector
    return f.v(f"shufflevector {V4} {a}, {V4} {a}, <3 x i32> <i32 0, i32 1, i32 2>")


def _pad3(f: _Fn, v: str) -> str:
    e = [f.v(f"extractelement {V3} {v}, i32 {i}") for i in range(3)]
    return _construct4(f, e + ["0.0"])


def _spill_matrix(f: _Fn, m: str) -> str:
    # "create temporary array / fill temporary array with values" (:1340-1350 and siblings)
    p = f.v(f"alloca {MAT}")
    for i in range(4):
        col = f.v(f"extractvalue {MAT} {m}, {i}")
        f.do(f"store {V4} {col}, ptr {f.v(f'getelementptr inbounds {MAT}, ptr {p}, i32 0, i32 {i}')}")
    return p


# The C helpers of spirv_compile.cpp:423-492, as IR. They are compiled by MSVC in the reference
# (/fp:precise: no contraction, no reassociation), so each is the scalar fmul/fadd chain its loop
# nest spells out.
HELPERS = r"""
define void @Float4x4TimesVec4(ptr %m, ptr %v, ptr %o) {
  ; :423-431  out[row] = 0; out[row] += m[col*4+row] * v[col]
@@mxv@@
  ret void
}
define void @Vec4TimesFloat4x4(ptr %m, ptr %v, ptr %o) {
  ; :443-451  out[row] = 0; out[row] += m[row*4+col] * v[col]
@@vxm@@
  ret void
}
define void @Float4x4TimesFloat4x4(ptr %a, ptr %b, ptr %o) {
  ; :463-473  out[x*4+y] = b[x*4+0]*a[0*4+y] + b[x*4+1]*a[1*4+y] + b[x*4+2]*a[2*4+y] + b[x*4+3]*a[3*4+y]
@@mxm@@
  ret void
}
define void @Float4x4TimesFloat(ptr %a, float %b, ptr %o) {
  ; :475-484
@@mxs@@
  ret void
}
define void @Float4x4Transpose(ptr %in, ptr %o) {
  ; :486-491  out[x*4+y] = in[y*4+x]
@@tr@@
  ret void
}
declare float @llvm.sqrt.f32(float)
declare float @llvm.sin.f32(float)
declare float @llvm.cos.f32(float)
declare float @llvm.pow.f32(float, float)
"""


def _helper_bodies() -> Dict[str, str]:
    def ld(f, p, i):
        return f.v(f"load float, ptr {f.v(f'getelementptr inbounds float, ptr {p}, i64 {i}')}, align 4")

    def st(f, p, i, val):
        f.do(f"store float {val}, ptr {f.v(f'getelementptr inbounds float, ptr {p}, i64 {i}')}, align 4")

    out = {}
    for name, idx in (("mxv", lambda r, c: c * 4 + r), ("vxm", lambda r, c: r * 4 + c)):
        f = _Fn()
        for r in range(4):
            acc = "0.0"
            for c in range(4):
                prod = f.v("fmul float %s, %s" % (ld(f, "%m", idx(r, c)), ld(f, "%v", c)))
                acc = f.v("fadd float %s, %s" % (acc, prod))
            st(f, "%o", r, acc)
        out[name] = "\n".join(f.lines)
    f = _Fn()
    for x in range(4):
        for y in range(4):
            acc = f.v("fmul float %s, %s" % (ld(f, "%b", x * 4), ld(f, "%a", y)))
            for k in range(1, 4):
                prod = f.v("fmul float %s, %s" % (ld(f, "%b", x * 4 + k), ld(f, "%a", k * 4 + y)))
                acc = f.v("fadd float %s, %s" % (acc, prod))
            st(f, "%o", x * 4 + y, acc)
    out["mxm"] = "\n".join(f.lines)
    f = _Fn()
    for i in range(16):
        st(f, "%o", i, f.v(f"fmul float {ld(f, '%a', i)}, %b"))
    out["mxs"] = "\n".join(f.lines)
    f = _Fn()
    for x in range(4):
        for y in range(4):
            st(f, "%o", x * 4 + y, ld(f, "%in", y * 4 + x))
    out["tr"] = "\n".join(f.lines)
    return out


def _op_body(op: str, f: _Fn) -> str:
    """IR for r = op(a, b, c, M, N); %a %b %c are <4 x float>, %M %N are [4 x <4 x float>]."""
    a, b, c = "%a", "%b", "%c"
    if op in ("fadd", "fsub", "fmul", "fdiv"):    # :1493-1512
        return f.v(f"{op} {V4} {a}, {b}")
    if op == "fneg":                              # :1513-1517; LLVM 6 CreateFNeg == fsub -0.0, x
        return f.v(f"fsub {V4} <float -0.0, float -0.0, float -0.0, float -0.0>, {a}")
    if op == "vts":                               # :1486-1492
        bx = f.v(f"extractelement {V4} {b}, i32 0")
        return f.v(f"fmul {V4} {a}, {_splat(f, bx)}")
    if op == "dot4":                              # :1747-1752
        d = _dot(f, a, b, 4)
        return _construct4(f, [d] * 4)
    if op == "dot3":
        d = _dot(f, _xyz(f, a), _xyz(f, b), 3)
        return _construct4(f, [d] * 4)
    if op in ("fmin", "fmax"):                    # :1545-1556
        pred = "olt" if op == "fmin" else "ogt"
        cmp = f.v(f"fcmp {pred} {V4} {a}, {b}")
        return f.v(f"select <4 x i1> {cmp}, {V4} {a}, {V4} {b}")
    if op == "fclamp":                            # :1557-1571 val=a lower=b upper=c
        up = f.v(f"select <4 x i1> {f.v(f'fcmp olt {V4} {a}, {c}')}, {V4} {a}, {V4} {c}")
        return f.v(f"select <4 x i1> {f.v(f'fcmp ogt {V4} {up}, {b}')}, {V4} {up}, {V4} {b}")
    if op == "fmix":                              # :1572-1590 x=a y=b a=c
        xmul = f.v(f"fsub {V4} {_splat(f, '1.0')}, {c}")
        return f.v(f"fadd {V4} {f.v(f'fmul {V4} {xmul}, {a}')}, {f.v(f'fmul {V4} {c}, {b}')}")
    ax = None
    if op in ("sqrt", "invsqrt", "sin", "cos"):
        ax = f.v(f"extractelement {V4} {a}, i32 0")
    if op in ("sqrt", "sin", "cos"):              # :1591-1614
        return _construct4(f, [f.v(f"call float @llvm.{op}.f32(float {ax})")] * 4)
    if op == "invsqrt":                           # :1615-1645 scalar branch
        s = f.v(f"call float @llvm.sqrt.f32(float {ax})")
        return _construct4(f, [f.v(f"fdiv float 1.0, {s}")] * 4)
    if op == "normalize3":                        # :1646-1658
        a3 = _xyz(f, a)
        ln = f.v(f"call float @llvm.sqrt.f32(float {_dot(f, a3, a3, 3)})")
        inv = _splat(f, f.v(f"fdiv float 1.0, {ln}"), 3)
        return _pad3(f, f.v(f"fmul {V3} {a3}, {inv}"))
    if op == "length3":                           # :1659-1667
        a3 = _xyz(f, a)
        return _construct4(f, [f.v(f"call float @llvm.sqrt.f32(float {_dot(f, a3, a3, 3)})")] * 4)
    if op == "reflect3":                          # :1699-1712
        i3, n3 = _xyz(f, a), _xyz(f, b)
        d2 = f.v(f"fmul float {_dot(f, i3, n3, 3)}, 2.0")
        return _pad3(f, f.v(f"fsub {V3} {i3}, {f.v(f'fmul {V3} {_splat(f, d2, 3)}, {n3}')}"))
    if op == "cross3":                            # :1668-1674 "TODO": the first operand
        return _pad3(f, _xyz(f, a))
    if op == "shuffle":                           # :1789-1833 equal sizes: shufflevector a, b, <3,4,1,6>
        return f.v(f"shufflevector {V4} {a}, {V4} {b}, <4 x i32> <i32 3, i32 4, i32 1, i32 6>")
    if op == "pow":                               # :1675-1698 vectors component-wise
        cur = "undef"
        for i in range(4):
            x = f.v(f"extractelement {V4} {a}, i32 {i}")
            y = f.v(f"extractelement {V4} {b}, i32 {i}")
            cur = f.v(f"insertelement {V4} {cur}, float {f.v(f'call float @llvm.pow.f32(float {x}, float {y})')}, i32 {i}")
        return cur
    if op in ("mxv", "vxm"):                      # :1322-1384
        ret = f.v(f"alloca {V4}")
        arr = _spill_matrix(f, "%M")
        # the helpers take `const float4 &`: the vector operand is passed by value in the IR call and
        # materialised by the x64 ABI as a pointer to a temporary
        vtmp = f.v(f"alloca {V4}")
        f.do(f"store {V4} {a}, ptr {vtmp}")
        fn = "Float4x4TimesVec4" if op == "mxv" else "Vec4TimesFloat4x4"
        f.do(f"call void @{fn}(ptr {arr}, ptr {vtmp}, ptr {ret})")
        return f.v(f"load {V4}, ptr {ret}")
    if op == "mxm":                               # :1385-1414, then the shader folds the columns
        ret = f.v(f"alloca {MAT}")
        ap, bp = _spill_matrix(f, "%M"), _spill_matrix(f, "%N")
        f.do(f"call void @Float4x4TimesFloat4x4(ptr {ap}, ptr {bp}, ptr {ret})")
        prod = f.v(f"load {MAT}, ptr {ret}")
        r = f.v(f"extractvalue {MAT} {prod}, 0")
        for i in (1, 2, 3):
            r = f.v(f"fadd {V4} {r}, {f.v(f'extractvalue {MAT} {prod}, {i}')}")
        return r
    if op in ("transpose", "minverse"):           # :1440-1463; MatrixInverse calls Float4x4Transpose (:1713-1740)
        ret = f.v(f"alloca {MAT}")
        mp = _spill_matrix(f, "%M")
        f.do(f"call void @Float4x4Transpose(ptr {mp}, ptr {ret})")
        return f.v(f"extractvalue {MAT} {f.v(f'load {MAT}, ptr {ret}')}, {1 if op == 'transpose' else 3}")
    if op == "mxs":                               # :1415-1439
        ret = f.v(f"alloca {MAT}")
        mp = _spill_matrix(f, "%M")
        ax = f.v(f"extractelement {V4} {a}, i32 0")
        f.do(f"call void @Float4x4TimesFloat(ptr {mp}, float {ax}, ptr {ret})")
        return f.v(f"extractvalue {MAT} {f.v(f'load {MAT}, ptr {ret}')}, 2")
    raise ValueError(op)


def module_ir(ops) -> str:
    helpers = HELPERS
    for k, body in _helper_bodies().items():
        helpers = helpers.replace("@@%s@@" % k, body)
    text = [helpers]
    for op in ops:
        f = _Fn()
        r = _op_body(op, f)
        body = "\n".join(f.lines)
        text.append(f"""
define void @unit_{op}(ptr %pa, ptr %pb, ptr %pc, ptr %pM, ptr %pN, ptr %pout) {{
  %a = load {V4}, ptr %pa, align 4
  %b = load {V4}, ptr %pb, align 4
  %c = load {V4}, ptr %pc, align 4
  %M = load {MAT}, ptr %pM, align 4
  %N = load {MAT}, ptr %pN, align 4
{body}
  store {V4} {r}, ptr %pout, align 4
  ret void
}}""")
    return "\n".join(text)


class Jit:
    """x86 code for the unit-op functions, generated by the LLVM bundled with llvmlite. No fast-math flags
    anywhere and the default (non-contracting) FP options: the reference's IRBuilder sets none (:676)."""

    def __init__(self, ops) -> None:
        import llvmlite.binding as llvm
        for init in ("initialize", "initialize_native_target", "initialize_native_asmprinter"):
            try:
                getattr(llvm, init)()
            except Exception:    # newer llvmlite initialises on import and deprecates these
                pass
        self.llvm = llvm
        mod = llvm.parse_assembly(module_ir(ops))
        mod.verify()
        target = llvm.Target.from_default_triple()
        self.tm = target.create_target_machine(opt=2)
        self.engine = llvm.create_mcjit_compiler(mod, self.tm)
        self.engine.finalize_object()
        self.ops = tuple(ops)
        self._fn: Dict[str, Callable] = {}
        proto = C.CFUNCTYPE(None, *([C.c_void_p] * 6))
        for op in ops:
            self._fn[op] = proto(self.engine.get_function_address(f"unit_{op}"))

    def run(self, op: str, a: np.ndarray, b: np.ndarray, c: np.ndarray, M: np.ndarray, N: np.ndarray) -> np.ndarray:
        """a, b, c: 4 floats; M, N: 16 floats, column-major as they sit in the UBO"""
        args = [np.ascontiguousarray(x, dtype=np.float32) for x in (a, b, c, M, N)]
        out = np.zeros(4, dtype=np.float32)
        self._fn[op](*[x.ctypes.data for x in args], out.ctypes.data)
        return out
