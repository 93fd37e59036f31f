"""Hand-assembled SPIR-V shaders for the BASELINE.json configs and for opcode-coverage tests.

Each function returns a numpy uint32 word array; entry points are always called "main".
Only the subset the reference's front end accepts is used: scalar-only
OpCompositeConstruct, no OpSelect/OpPhi, UBO members whose LLVM natural layout coincides with
std140 (mat4 / vec4).
"""
from __future__ import annotations

import numpy as np

from .spvasm import FRAGMENT, GLSL, SC, VERTEX, BuiltIn, Dec, Module, Op


def _main(m: Module):
    void = m.t_void()
    f, _ = m.begin_function(void, m.t_func(void))
    m.label()
    return f


def _finish(m: Module, model: int, f: int, iface) -> np.ndarray:
    m.ret()
    m.end_function()
    m.entry_point(model, f, "main", iface)
    return m.words()


# ------------------------------------------------------------------ C1: passthrough
def vs_passthrough() -> np.ndarray:
    """in vec4 pos@0, vec4 col@1; gl_Position = pos; out vec4 col@0."""
    m = Module()
    v4 = m.t_fvec(4)
    pos = m.input(v4, 0, "pos")
    col = m.input(v4, 1, "col")
    ocol = m.output(v4, 0, "ocol")
    gl = m.per_vertex_out()
    f = _main(m)
    p = m.load(v4, pos)
    m.store(m.access(SC.Output, v4, gl, m.const_i(0)), p)
    m.store(ocol, m.load(v4, col))
    return _finish(m, VERTEX, f, [pos, col, ocol, gl])


def fs_color() -> np.ndarray:
    """in vec4 col@0; out vec4 o@0 = col."""
    m = Module()
    v4 = m.t_fvec(4)
    col = m.input(v4, 0, "col")
    o = m.output(v4, 0, "o")
    f = _main(m)
    m.store(o, m.load(v4, col))
    return _finish(m, FRAGMENT, f, [col, o])


def fs_color_kill(threshold: float = 0.5, in_callee: bool = False) -> np.ndarray:
    """in vec4 col@0; if(col.x < threshold) discard; out vec4 o@0 = col.   (extended mode: OpKill)
    in_callee: the discard sits in a helper function the entry point calls (the kill must end the invocation,
    not just the helper)."""
    m = Module()
    v4, fl, bt, void = m.t_fvec(4), m.t_float(), m.t_bool(), m.t_void()
    col = m.input(v4, 0, "col")
    o = m.output(v4, 0, "o")

    def test_and_kill(c):
        kill, merge = m.new_id(), m.new_id()
        cond = m.inst(Op.FOrdLessThan, bt, m.extract(fl, c, 0), m.const_f(threshold))
        m.stmt(Op.SelectionMerge, merge, 0)
        m.stmt(Op.BranchConditional, cond, kill, merge)
        m.label(kill)
        m.stmt(Op.Kill)
        m.label(merge)

    helper = None
    if in_callee:
        helper, (hp,) = m.begin_function(void, m.t_func(void, v4), [v4])
        m.label()
        test_and_kill(hp)
        m.ret()
        m.end_function()
    f = _main(m)
    c = m.load(v4, col)
    if in_callee:
        m.inst(Op.FunctionCall, void, helper, c)
    else:
        test_and_kill(c)
    m.store(o, c)
    return _finish(m, FRAGMENT, f, [col, o])


# ------------------------------------------------------------------ C2: textured cube
def vs_mvp_uv() -> np.ndarray:
    """UBO{mat4 mvp}@(0,0); in vec4 pos@0, vec2 uv@1; gl_Position = mvp*pos; out vec2 uv@0."""
    m = Module()
    v4, v2, mat4 = m.t_fvec(4), m.t_fvec(2), m.t_mat(4)
    st = m.t_struct(mat4, tag="UBO")
    m.member_decorate(st, 0, Dec.ColMajor)
    m.member_decorate(st, 0, Dec.Offset, 0)
    m.member_decorate(st, 0, Dec.MatrixStride, 16)
    ubo = m.ubo(st, 0, 0, "ubo")
    pos = m.input(v4, 0, "pos")
    uv = m.input(v2, 1, "uv")
    ouv = m.output(v2, 0, "ouv")
    gl = m.per_vertex_out()
    f = _main(m)
    mvp = m.load(mat4, m.access(SC.Uniform, mat4, ubo, m.const_i(0)))
    r = m.inst(Op.MatrixTimesVector, v4, mvp, m.load(v4, pos))
    m.store(m.access(SC.Output, v4, gl, m.const_i(0)), r)
    m.store(ouv, m.load(v2, uv))
    return _finish(m, VERTEX, f, [pos, uv, ouv, gl])


def fs_texture(explicit_lod: bool = False) -> np.ndarray:
    """sampler2D tex@(0,1); in vec2 uv@0; out = texture(tex, uv) — or textureLod(tex, uv, 2.0) (extended mode;
    the level is ignored like everything else about the sampler: mip 0)."""
    m = Module()
    v4, v2 = m.t_fvec(4), m.t_fvec(2)
    tex = m.sampler2d(0, 1, "tex")
    uv = m.input(v2, 0, "uv")
    o = m.output(v4, 0, "o")
    f = _main(m)
    s = m.load(m.t_sampled_image(), tex)
    if explicit_lod:
        c = m.inst(Op.ImageSampleExplicitLod, v4, s, m.load(v2, uv), 0x2, m.const_f(2.0))    # image operands: Lod
    else:
        c = m.inst(Op.ImageSampleImplicitLod, v4, s, m.load(v2, uv))
    m.store(o, c)
    return _finish(m, FRAGMENT, f, [uv, o])


# ------------------------------------------------------------------ C3 / C5: lit mesh
def vs_lit(with_uv: bool = False) -> np.ndarray:
    """UBO{mat4 mvp; vec4 light; vec4 albedo; vec4 ambient}@(0,0);
    in vec3 pos@0, vec3 nrm@1, vec2 uv@2.
    gl_Position = mvp*vec4(pos,1); col = albedo*max(dot(nrm,light.xyz),0) + ambient -> out@0
    (+ uv -> out@1)."""
    m = Module()
    fl, v4, v3, v2, mat4 = m.t_float(), m.t_fvec(4), m.t_fvec(3), m.t_fvec(2), m.t_mat(4)
    st = m.t_struct(mat4, v4, v4, v4, tag="UBO")
    m.member_decorate(st, 0, Dec.ColMajor)
    m.member_decorate(st, 0, Dec.Offset, 0)
    m.member_decorate(st, 0, Dec.MatrixStride, 16)
    for i, off in ((1, 64), (2, 80), (3, 96)):
        m.member_decorate(st, i, Dec.Offset, off)
    ubo = m.ubo(st, 0, 0, "ubo")
    pos = m.input(v3, 0, "pos")
    nrm = m.input(v3, 1, "nrm")
    uv = m.input(v2, 2, "uv") if with_uv else None
    ocol = m.output(v4, 0, "ocol")
    ouv = m.output(v2, 1, "ouv") if with_uv else None
    gl = m.per_vertex_out()
    f = _main(m)
    p = m.load(v3, pos)
    p4 = m.construct(v4, m.extract(fl, p, 0), m.extract(fl, p, 1), m.extract(fl, p, 2), m.const_f(1.0))
    mvp = m.load(mat4, m.access(SC.Uniform, mat4, ubo, m.const_i(0)))
    m.store(m.access(SC.Output, v4, gl, m.const_i(0)), m.inst(Op.MatrixTimesVector, v4, mvp, p4))
    light = m.load(v4, m.access(SC.Uniform, v4, ubo, m.const_i(1)))
    l3 = m.shuffle(v3, light, light, 0, 1, 2)
    ndl = m.inst(Op.Dot, fl, m.load(v3, nrm), l3)
    ndl = m.ext(fl, GLSL.FMax, ndl, m.const_f(0.0))
    albedo = m.load(v4, m.access(SC.Uniform, v4, ubo, m.const_i(2)))
    ambient = m.load(v4, m.access(SC.Uniform, v4, ubo, m.const_i(3)))
    lit = m.inst(Op.VectorTimesScalar, v4, albedo, ndl)
    m.store(ocol, m.inst(Op.FAdd, v4, lit, ambient))
    iface = [pos, nrm, ocol, gl]
    if with_uv:
        m.store(ouv, m.load(v2, uv))
        iface += [uv, ouv]
    return _finish(m, VERTEX, f, iface)


def fs_lit_tex() -> np.ndarray:
    """sampler2D tex@(0,1); in vec4 col@0, vec2 uv@1; out = texture(tex,uv) * col."""
    m = Module()
    v4, v2 = m.t_fvec(4), m.t_fvec(2)
    tex = m.sampler2d(0, 1, "tex")
    col = m.input(v4, 0, "col")
    uv = m.input(v2, 1, "uv")
    o = m.output(v4, 0, "o")
    f = _main(m)
    s = m.load(m.t_sampled_image(), tex)
    t = m.inst(Op.ImageSampleImplicitLod, v4, s, m.load(v2, uv))
    m.store(o, m.inst(Op.FMul, v4, t, m.load(v4, col)))
    return _finish(m, FRAGMENT, f, [col, uv, o])


# ------------------------------------------------------------------ opcode coverage
def vs_kitchen_sink() -> np.ndarray:
    """Exercises most of Appendix B on the vertex side.

    push constants {vec4 k} (offset 0); UBO{mat4 a; mat4 b; vec4 s}@(1,2);
    in vec4 pos@0, vec3 nrm@1, float w@2, int flag@3 (R32_SINT), gl_VertexIndex.
    out@0 vec4, out@1 vec3, out@2 float, out@3 int (flat), out@4 vec2, out@5..8 mat4.
    Uses: function call with pointer param, loop (SLessThan/IAdd/ConvertSToF), branch, FMix, FClamp,
    FMin, FMax, Normalize, Length, Reflect, Sqrt, InverseSqrt, FNegate, FDiv, FSub, VectorShuffle,
    VectorTimesMatrix, MatrixTimesMatrix, MatrixTimesScalar, Transpose, MatrixInverse(=transpose),
    Cross(=arg0), IMul, BitwiseAnd, ShiftLeftLogical(=shift right), IEqual, FOrdLessThan(Equal),
    FOrdGreaterThan, CompositeExtract on matrix, CompositeConstruct of matrix.
    """
    m = Module()
    fl, it = m.t_float(), m.t_int(1)
    v4, v3, v2, mat4 = m.t_fvec(4), m.t_fvec(3), m.t_fvec(2), m.t_mat(4)
    bl = m.t_bool()
    pc_t = m.t_struct(v4, tag="PC")
    m.member_decorate(pc_t, 0, Dec.Offset, 0)
    pc = m.push_constants(pc_t, "pc")
    ubo_t = m.t_struct(mat4, mat4, v4, tag="UBO")
    for i, off in ((0, 0), (1, 64), (2, 128)):
        m.member_decorate(ubo_t, i, Dec.Offset, off)
    ubo = m.ubo(ubo_t, 1, 2, "ubo")
    pos = m.input(v4, 0, "pos")
    nrm = m.input(v3, 1, "nrm")
    win = m.input(fl, 2, "w")
    flag = m.input(it, 3, "flag")
    vidx = m.builtin_input(it, BuiltIn.VertexIndex, "gl_VertexIndex")
    o0 = m.output(v4, 0)
    o1 = m.output(v3, 1)
    o2 = m.output(fl, 2)
    o3 = m.output(it, 3)
    o4 = m.output(v2, 4)
    o5 = m.output(mat4, 5)
    gl = m.per_vertex_out()

    void = m.t_void()
    # float helper(inout float acc, vec3 n): acc = acc + length(n); return sqrt(acc)
    pfl = m.t_ptr(SC.Function, fl)
    hf, (h_acc, h_n) = m.begin_function(fl, m.t_func(fl, pfl, v3), [pfl, v3])
    m.label()
    a = m.load(fl, h_acc)
    a2 = m.inst(Op.FAdd, fl, a, m.ext(fl, GLSL.Length, h_n))
    m.store(h_acc, a2)
    m.stmt(Op.ReturnValue, m.ext(fl, GLSL.Sqrt, a2))
    m.end_function()

    f, _ = m.begin_function(void, m.t_func(void))
    m.label()
    acc = m.local(fl, m.const_f(0.25))
    i_var = m.local(it, m.const_i(0))
    outsel = m.local(fl)
    p = m.load(v4, pos)
    n = m.load(v3, nrm)
    k = m.load(v4, m.access(SC.PushConstant, v4, pc, m.const_i(0)))
    ma = m.load(mat4, m.access(SC.Uniform, mat4, ubo, m.const_i(0)))
    mb = m.load(mat4, m.access(SC.Uniform, mat4, ubo, m.const_i(1)))
    s = m.load(v4, m.access(SC.Uniform, v4, ubo, m.const_i(2)))
    # scalar member read through a deeper access chain: s.y
    sy = m.load(fl, m.access(SC.Uniform, fl, ubo, m.const_i(2), m.const_i(1)))

    mm = m.inst(Op.MatrixTimesMatrix, mat4, ma, mb)
    mt = m.inst(Op.Transpose, mat4, mm)
    ms = m.inst(Op.MatrixTimesScalar, mat4, mt, sy)
    mi = m.ext(mat4, GLSL.MatrixInverse, ms)
    pv = m.inst(Op.MatrixTimesVector, v4, ma, p)
    vp = m.inst(Op.VectorTimesMatrix, v4, p, mb)
    col1 = m.extract(v4, mi, 1)
    e23 = m.extract(fl, mi, 2, 3)
    rebuilt = m.construct(mat4, pv, vp, col1, k)

    # loop: for (i = 0; i < 3; i++) acc += float(i) * w
    head, body, cont, merge = m.new_id(), m.new_id(), m.new_id(), m.new_id()
    m.stmt(Op.Branch, head)
    m.label(head)
    iv = m.load(it, i_var)
    cond = m.inst(Op.SLessThan, bl, iv, m.const_i(3))
    m.stmt(Op.LoopMerge, merge, cont, 0)
    m.stmt(Op.BranchConditional, cond, body, merge)
    m.label(body)
    fi = m.inst(Op.ConvertSToF, fl, iv)
    m.store(acc, m.inst(Op.FAdd, fl, m.load(fl, acc), m.inst(Op.FMul, fl, fi, m.load(fl, win))))
    m.stmt(Op.Branch, cont)
    m.label(cont)
    m.store(i_var, m.inst(Op.IAdd, it, m.load(it, i_var), m.const_i(1)))
    m.stmt(Op.Branch, head)
    m.label(merge)

    hres = m.inst(Op.FunctionCall, fl, hf, acc, n)

    nn = m.ext(v3, GLSL.Normalize, n)
    refl = m.ext(v3, GLSL.Reflect, m.shuffle(v3, p, p, 0, 1, 2), nn)
    crs = m.ext(v3, GLSL.Cross, refl, nn)
    mixv = m.ext(v4, GLSL.FMix, pv, vp, s)
    clv = m.ext(v4, GLSL.FClamp, mixv, m.const_fvec(-0.5, -0.5, -0.5, -0.5), m.const_fvec(2, 2, 2, 2))
    mn = m.ext(fl, GLSL.FMin, hres, e23)
    mx = m.ext(fl, GLSL.FMax, hres, e23)
    isq = m.ext(fl, GLSL.InverseSqrt, m.inst(Op.FAdd, fl, mx, m.const_f(1.5)))
    neg = m.inst(Op.FNegate, fl, mn)
    dv = m.inst(Op.FDiv, fl, neg, m.inst(Op.FAdd, fl, isq, m.const_f(3.0)))
    sb = m.inst(Op.FSub, fl, dv, sy)

    # integer ops + branch: flag2 = ((flag * 3) & 0xff) >> 1 [the reference emits lshr for SHL]
    fg = m.load(it, flag)
    f2 = m.inst(Op.IMul, it, fg, m.const_i(3))
    f3 = m.inst(Op.BitwiseAnd, it, f2, m.const_i(0xFF))
    f4 = m.inst(Op.ShiftLeftLogical, it, f3, m.const_i(1))
    f5 = m.inst(Op.IAdd, it, f4, m.load(it, vidx))
    m.store(outsel, sb)
    is7 = m.inst(Op.IEqual, bl, fg, m.const_i(7))
    t_lbl, e_lbl, j_lbl = m.new_id(), m.new_id(), m.new_id()
    m.stmt(Op.SelectionMerge, j_lbl, 0)
    m.stmt(Op.BranchConditional, is7, t_lbl, e_lbl)
    m.label(t_lbl)
    m.store(outsel, m.inst(Op.FMul, fl, sb, m.const_f(2.0)))
    m.stmt(Op.Branch, j_lbl)
    m.label(e_lbl)
    lt = m.inst(Op.FOrdLessThan, bl, sb, m.const_f(0.0))
    t2, j2 = m.new_id(), m.new_id()
    m.stmt(Op.SelectionMerge, j2, 0)
    m.stmt(Op.BranchConditional, lt, t2, j2)
    m.label(t2)
    m.store(outsel, m.inst(Op.FAdd, fl, sb, m.const_f(10.0)))
    m.stmt(Op.Branch, j2)
    m.label(j2)
    m.stmt(Op.Branch, j_lbl)
    m.label(j_lbl)
    le = m.inst(Op.FOrdLessThanEqual, bl, m.load(fl, outsel), m.const_f(100.0))
    t3, j3 = m.new_id(), m.new_id()
    m.stmt(Op.SelectionMerge, j3, 0)
    m.stmt(Op.BranchConditional, le, t3, j3)
    m.label(t3)
    gt = m.inst(Op.FOrdGreaterThan, bl, m.load(fl, outsel), m.const_f(-100.0))
    t4, j4 = m.new_id(), m.new_id()
    m.stmt(Op.SelectionMerge, j4, 0)
    m.stmt(Op.BranchConditional, gt, t4, j4)
    m.label(t4)
    m.store(outsel, m.inst(Op.FAdd, fl, m.load(fl, outsel), m.const_f(0.125)))
    m.stmt(Op.Branch, j4)
    m.label(j4)
    m.stmt(Op.Branch, j3)
    m.label(j3)

    m.store(m.access(SC.Output, v4, gl, m.const_i(0)), p)
    m.store(o0, clv)
    m.store(o1, crs)
    m.store(o2, m.load(fl, outsel))
    m.store(o3, f5)
    m.store(o4, m.shuffle(v2, k, s, 1, 6))
    m.store(o5, rebuilt)
    return _finish(m, VERTEX, f, [pos, nrm, win, flag, vidx, o0, o1, o2, o3, o4, o5, gl])


def fs_kitchen_sink() -> np.ndarray:
    """in@0 vec4, @1 vec3, @2 float, @3 int (flat), @4 vec2, @5 mat4 (array of 4 slots);
    push constants {vec4 k}; out = clamp-free mix of everything (kept inside [0,1] by the test)."""
    m = Module()
    fl, it = m.t_float(), m.t_int(1)
    v4, v3, v2, mat4 = m.t_fvec(4), m.t_fvec(3), m.t_fvec(2), m.t_mat(4)
    pc_t = m.t_struct(v4, tag="PC")
    m.member_decorate(pc_t, 0, Dec.Offset, 0)
    pc = m.push_constants(pc_t, "pc")
    i0 = m.input(v4, 0)
    i1 = m.input(v3, 1)
    i2 = m.input(fl, 2)
    i3 = m.input(it, 3)
    m.decorate(i3, Dec.Flat)
    i4 = m.input(v2, 4)
    i5 = m.input(mat4, 5)
    o = m.output(v4, 0)
    f = _main(m)
    a = m.load(v4, i0)
    b = m.load(v3, i1)
    c = m.load(fl, i2)
    d = m.inst(Op.ConvertSToF, fl, m.load(it, i3))
    e = m.load(v2, i4)
    mt = m.load(mat4, i5)
    k = m.load(v4, m.access(SC.PushConstant, v4, pc, m.const_i(0)))
    mv = m.inst(Op.MatrixTimesVector, v4, mt, k)
    r = m.inst(Op.FAdd, v4, a, mv)
    r = m.inst(Op.VectorTimesScalar, v4, r, m.const_f(0.03125))
    bx = m.extract(fl, b, 0)
    ex = m.extract(fl, e, 1)
    t = m.inst(Op.FMul, fl, m.inst(Op.FAdd, fl, m.inst(Op.FAdd, fl, bx, ex), c), m.const_f(0.015625))
    t = m.inst(Op.FAdd, fl, t, m.inst(Op.FMul, fl, d, m.const_f(0.0009765625)))
    tv = m.construct(v4, t, t, t, t)
    m.store(o, m.inst(Op.FAdd, v4, r, tv))
    return _finish(m, FRAGMENT, f, [i0, i1, i2, i3, i4, i5, o])


def fs_cube() -> np.ndarray:
    """samplerCube tex@(0,1); in vec3 dir@0; out = texture(tex, dir)."""
    m = Module()
    v4, v3 = m.t_fvec(4), m.t_fvec(3)
    tex = m.sampler2d(0, 1, "tex", dim=3)
    d = m.input(v3, 0, "dir")
    o = m.output(v4, 0, "o")
    f = _main(m)
    s = m.load(m.t_sampled_image(3), tex)
    m.store(o, m.inst(Op.ImageSampleImplicitLod, v4, s, m.load(v3, d)))
    return _finish(m, FRAGMENT, f, [d, o])


def vs_pos_dir() -> np.ndarray:
    """in vec4 pos@0, vec3 dir@1; gl_Position = pos; out vec3 dir@0."""
    m = Module()
    v4, v3 = m.t_fvec(4), m.t_fvec(3)
    pos = m.input(v4, 0)
    d = m.input(v3, 1)
    od = m.output(v3, 0)
    gl = m.per_vertex_out()
    f = _main(m)
    m.store(m.access(SC.Output, v4, gl, m.const_i(0)), m.load(v4, pos))
    m.store(od, m.load(v3, d))
    return _finish(m, VERTEX, f, [pos, d, od, gl])


# ------------------------------------------------------------------ single-op shaders (known-answer tests)
UNIT_OPS = ("fadd", "fsub", "fmul", "fdiv", "fneg", "vts", "dot4", "dot3", "fmin", "fmax", "fclamp", "fmix",
            "sqrt", "invsqrt", "normalize3", "length3", "reflect3", "cross3", "shuffle", "mxv", "vxm", "mxm",
            "transpose", "mxs", "minverse", "sin", "cos", "pow")


# memory-model cases of the reference's subset that are not single arithmetic opcodes: run-time indices into
# function-local arrays and vectors (OpAccessChain -> GEP, spirv_compile.cpp:1301-1318)
MEM_UNIT_OPS = ("dynidx",)


# opcodes outside the reference's subset (SURVEY.md Appendix B "Not supported"), accepted only when the
# "extended_spirv" option is on
EXT_UNIT_OPS = ("select", "fge", "feq", "fne", "isub_bitcast", "ftos", "fabs", "floor", "fract",
                "roundeven", "trunc", "ceil", "fsign", "radians", "degrees", "step", "smoothstep", "fma",
                "distance3", "faceforward3", "refract3", "int_minmax", "uint_minmax", "int_abs_sign", "phi_loop",
                "phi_swap", "int_divmod", "uint_divmod", "shifts_bits", "ucvt", "int_cmp", "logic", "isnan_inf",
                "switch_phi", "consts_copy", "composite_insert", "vec_dynamic", "frem_fmod", "any_all", "bit_ops",
                "nminmax", "exp_log", "tan_hyp", "atan_asin", "bitfield", "determinant", "funord", "inv_hyp")
# of those, the ones built on transcendental functions: libm in the oracle, the special-function unit on the GPU,
# compared under the 1-LSB colour bar like sin / cos / pow
APPROX_EXT_OPS = ("exp_log", "tan_hyp", "atan_asin", "inv_hyp")


def vs_unit(op: str) -> np.ndarray:
    """in vec4 a@0, b@1, c@2; UBO{mat4 m; mat4 n}@(0,0); gl_Position = a; out vec4 r@0 = op(a,b,c,m,n)."""
    m = Module()
    fl, v4, v3, mat4 = m.t_float(), m.t_fvec(4), m.t_fvec(3), m.t_mat(4)
    st = m.t_struct(mat4, mat4, tag="UBO")
    m.member_decorate(st, 0, Dec.Offset, 0)
    m.member_decorate(st, 1, Dec.Offset, 64)
    ubo = m.ubo(st, 0, 0, "ubo")
    ia, ib, ic = m.input(v4, 0, "a"), m.input(v4, 1, "b"), m.input(v4, 2, "c")
    out = m.output(v4, 0, "r")
    gl = m.per_vertex_out()
    f = _main(m)
    a, b, c = m.load(v4, ia), m.load(v4, ib), m.load(v4, ic)
    M = m.load(mat4, m.access(SC.Uniform, mat4, ubo, m.const_i(0)))
    N = m.load(mat4, m.access(SC.Uniform, mat4, ubo, m.const_i(1)))
    a3, b3 = m.shuffle(v3, a, a, 0, 1, 2), m.shuffle(v3, b, b, 0, 1, 2)
    ax, bx = m.extract(fl, a, 0), m.extract(fl, b, 0)
    extra_iface = []

    def splat(s):
        return m.construct(v4, s, s, s, s)

    def pad3(v):
        return m.construct(v4, m.extract(fl, v, 0), m.extract(fl, v, 1), m.extract(fl, v, 2), m.const_f(0.0))

    def col_sum(mat):
        # fold a matrix into a vec4: sum of its columns, left to right
        r = m.extract(v4, mat, 0)
        for i in (1, 2, 3):
            r = m.inst(Op.FAdd, v4, r, m.extract(v4, mat, i))
        return r

    if op == "fadd":
        r = m.inst(Op.FAdd, v4, a, b)
    elif op == "fsub":
        r = m.inst(Op.FSub, v4, a, b)
    elif op == "fmul":
        r = m.inst(Op.FMul, v4, a, b)
    elif op == "fdiv":
        r = m.inst(Op.FDiv, v4, a, b)
    elif op == "fneg":
        r = m.inst(Op.FNegate, v4, a)
    elif op == "vts":
        r = m.inst(Op.VectorTimesScalar, v4, a, bx)
    elif op == "dot4":
        r = splat(m.inst(Op.Dot, fl, a, b))
    elif op == "dot3":
        r = splat(m.inst(Op.Dot, fl, a3, b3))
    elif op == "fmin":
        r = m.ext(v4, GLSL.FMin, a, b)
    elif op == "fmax":
        r = m.ext(v4, GLSL.FMax, a, b)
    elif op == "fclamp":
        r = m.ext(v4, GLSL.FClamp, a, b, c)
    elif op == "fmix":
        r = m.ext(v4, GLSL.FMix, a, b, c)
    elif op == "sqrt":
        r = splat(m.ext(fl, GLSL.Sqrt, ax))
    elif op == "invsqrt":
        r = splat(m.ext(fl, GLSL.InverseSqrt, ax))
    elif op == "normalize3":
        r = pad3(m.ext(v3, GLSL.Normalize, a3))
    elif op == "length3":
        r = splat(m.ext(fl, GLSL.Length, a3))
    elif op == "reflect3":
        r = pad3(m.ext(v3, GLSL.Reflect, a3, b3))
    elif op == "cross3":
        r = pad3(m.ext(v3, GLSL.Cross, a3, b3))
    elif op == "shuffle":
        r = m.shuffle(v4, a, b, 3, 4, 1, 6)
    elif op == "mxv":
        r = m.inst(Op.MatrixTimesVector, v4, M, a)
    elif op == "vxm":
        r = m.inst(Op.VectorTimesMatrix, v4, a, M)
    elif op == "mxm":
        r = col_sum(m.inst(Op.MatrixTimesMatrix, mat4, M, N))
    elif op == "transpose":
        r = m.extract(v4, m.inst(Op.Transpose, mat4, M), 1)
    elif op == "mxs":
        r = m.extract(v4, m.inst(Op.MatrixTimesScalar, mat4, M, ax), 2)
    elif op == "minverse":
        r = m.extract(v4, m.ext(mat4, GLSL.MatrixInverse, M), 3)
    elif op == "sin":
        r = splat(m.ext(fl, GLSL.Sin, ax))
    elif op == "cos":
        r = splat(m.ext(fl, GLSL.Cos, ax))
    elif op == "pow":
        r = m.ext(v4, GLSL.Pow, a, b)
    elif op in ("select", "fge", "feq", "fne"):
        bv4 = m.t_vec(m.t_bool(), 4)
        cmp = {"select": Op.FOrdLessThan, "fge": Op.FOrdGreaterThanEqual, "feq": Op.FOrdEqual,
               "fne": Op.FOrdNotEqual}[op]
        r = m.inst(Op.Select, v4, m.inst(cmp, bv4, a, b), a, c)
    elif op == "isub_bitcast":
        iv4 = m.t_vec(m.t_int(1), 4)
        r = m.inst(Op.Bitcast, v4, m.inst(Op.ISub, iv4, m.inst(Op.Bitcast, iv4, a), m.inst(Op.Bitcast, iv4, b)))
    elif op == "ftos":
        iv4 = m.t_vec(m.t_int(1), 4)
        r = m.inst(Op.ConvertSToF, v4, m.inst(Op.ConvertFToS, iv4, a))
    elif op == "dynidx":
        # i = gl_VertexIndex & 3, j = (gl_VertexIndex * 3) & 3
        # vec4 arr[4] = {a, b, c, a + b};  float fa[4] = {a.x, b.y, c.z, a.w};  vec4 vv = b;
        # fa[i] = c.x;  r = arr[i] + vec4(fa[1], fa[j], vv[i], fa[1])
        it = m.t_int(1)
        vidx = m.builtin_input(it, BuiltIn.VertexIndex, "gl_VertexIndex")
        extra_iface.append(vidx)
        vi = m.load(it, vidx)
        i = m.inst(Op.BitwiseAnd, it, vi, m.const_i(3))
        j = m.inst(Op.BitwiseAnd, it, m.inst(Op.IMul, it, vi, m.const_i(3)), m.const_i(3))
        arr = m.local(m.t_array(v4, 4))
        for k, val in enumerate((a, b, c, m.inst(Op.FAdd, v4, a, b))):
            m.store(m.access(SC.Function, v4, arr, m.const_i(k)), val)
        fa = m.local(m.t_array(fl, 4))
        for k, (src, comp) in enumerate(((a, 0), (b, 1), (c, 2), (a, 3))):
            m.store(m.access(SC.Function, fl, fa, m.const_i(k)), m.extract(fl, src, comp))
        vv = m.local(v4)
        m.store(vv, b)
        m.store(m.access(SC.Function, fl, fa, i), m.extract(fl, c, 0))
        x = m.load(v4, m.access(SC.Function, v4, arr, i))
        s1 = m.load(fl, m.access(SC.Function, fl, fa, m.const_i(1)))
        t1 = m.load(fl, m.access(SC.Function, fl, fa, j))
        comp = m.load(fl, m.access(SC.Function, fl, vv, i))
        r = m.inst(Op.FAdd, v4, x, m.construct(v4, s1, t1, comp, s1))
    elif op in ("roundeven", "trunc", "ceil", "fsign", "radians", "degrees"):
        code = {"roundeven": GLSL.RoundEven, "trunc": GLSL.Trunc, "ceil": GLSL.Ceil, "fsign": GLSL.FSign,
                "radians": GLSL.Radians, "degrees": GLSL.Degrees}[op]
        arg = a
        if op in ("roundeven", "trunc", "ceil"):
            arg = m.inst(Op.FMul, v4, a, m.const_fvec(3.5, 2.5, 7.25, 0.5))    # ties and magnitudes above 1
        r = m.ext(v4, code, arg)
        if op in ("roundeven", "trunc", "ceil", "degrees"):
            r = m.inst(Op.FMul, v4, r, m.const_fvec(0.125, 0.125, 0.125, 0.125))
            if op == "degrees":
                r = m.inst(Op.FMul, v4, r, m.const_fvec(0.125, 0.125, 0.125, 0.125))
    elif op == "step":
        r = m.ext(v4, GLSL.Step, a, b)
    elif op == "smoothstep":
        r = m.ext(v4, GLSL.SmoothStep, a, b, c)
    elif op == "fma":
        r = m.ext(v4, GLSL.Fma, a, b, c)
    elif op == "distance3":
        r = splat(m.ext(fl, GLSL.Distance, a3, b3))
    elif op == "faceforward3":
        c3 = m.shuffle(v3, c, c, 0, 1, 2)
        r = pad3(m.ext(v3, GLSL.FaceForward, a3, b3, m.inst(Op.FSub, v3, c3, a3)))
    elif op == "refract3":
        r = pad3(m.ext(v3, GLSL.Refract, m.ext(v3, GLSL.Normalize, a3), m.ext(v3, GLSL.Normalize, b3),
                       m.extract(fl, c, 0)))
    elif op in ("int_minmax", "uint_minmax", "int_abs_sign"):
        signed = op != "uint_minmax"
        it = m.t_int(1 if signed else 0)
        iv4 = m.t_vec(it, 4)
        sv4 = m.t_vec(m.t_int(1), 4)
        # integers from the floats: (int)(x * 64) - 24 (negative values included), reinterpreted when unsigned
        def ints(x):
            k = m.inst(Op.ConvertFToS, sv4, m.inst(Op.FMul, v4, x, m.const_fvec(64.0, 64.0, 64.0, 64.0)))
            k = m.inst(Op.ISub, sv4, k, m.inst(Op.ConvertFToS, sv4, m.const_fvec(24.0, 24.0, 24.0, 24.0)))
            return k if signed else m.inst(Op.Bitcast, iv4, k)
        ia_, ib_, ic_ = ints(a), ints(b), ints(c)
        if op == "int_abs_sign":
            x = m.inst(Op.IAdd, iv4, m.ext(iv4, GLSL.SAbs, ia_), m.ext(iv4, GLSL.SSign, ib_))
        else:
            lo = m.ext(iv4, GLSL.SMin if signed else GLSL.UMin, ia_, ib_)
            hi = m.ext(iv4, GLSL.SMax if signed else GLSL.UMax, ia_, ib_)
            cl = m.ext(iv4, GLSL.SClamp if signed else GLSL.UClamp, ic_, lo, hi)
            x = m.inst(Op.IAdd, iv4, m.inst(Op.IAdd, iv4, lo, hi), cl)
        x = m.inst(Op.BitwiseAnd, iv4, x, m.inst(Op.Bitcast, iv4, m.inst(Op.ConvertFToS, sv4, m.const_fvec(255.0, 255.0, 255.0, 255.0))))
        xs = x if signed else m.inst(Op.Bitcast, sv4, x)
        r = m.inst(Op.FMul, v4, m.inst(Op.ConvertSToF, v4, xs), m.const_fvec(1 / 256.0, 1 / 256.0, 1 / 256.0, 1 / 256.0))
    elif op in ("phi_loop", "phi_swap"):
        # for(i = 0, acc = a, other = b; i < n; i++) { acc = acc * 0.5 + other * c; [swap: (acc, other) = (other, acc)] }
        # written in SSA form with OpPhi, as a real compiler's output has it (n = 3 + (int(a.x * 8) & 3))
        it = m.t_int(1)
        bt = m.t_bool()
        n = m.inst(Op.IAdd, it, m.const_i(3), m.inst(Op.BitwiseAnd, it, m.inst(Op.ConvertFToS, it,
                   m.inst(Op.FMul, fl, ax, m.const_f(8.0))), m.const_i(3)))
        entry = m.cur_label
        head, body, cont, merge = m.new_id(), m.new_id(), m.new_id(), m.new_id()
        m.stmt(Op.Branch, head)
        m.label(head)
        i_phi, acc_phi, oth_phi = m.new_id(), m.new_id(), m.new_id()
        i_next, acc_next, oth_next = m.new_id(), m.new_id(), m.new_id()
        m.raw(Op.Phi, it, i_phi, m.const_i(0), entry, i_next, cont)
        if op == "phi_swap":
            # acc takes the OLD value of the other phi of the same block: the copies of an edge are parallel
            m.raw(Op.Phi, v4, acc_phi, a, entry, oth_phi, cont)
            m.raw(Op.Phi, v4, oth_phi, b, entry, oth_next, cont)
        else:
            m.raw(Op.Phi, v4, acc_phi, a, entry, acc_next, cont)
            m.raw(Op.Phi, v4, oth_phi, b, entry, oth_next, cont)
        cond = m.inst(Op.SLessThan, bt, i_phi, n)
        m.stmt(Op.LoopMerge, merge, cont, 0)
        m.stmt(Op.BranchConditional, cond, body, merge)
        m.label(body)
        half = m.inst(Op.FMul, v4, acc_phi, m.const_fvec(0.5, 0.5, 0.5, 0.5))
        val = m.inst(Op.FAdd, v4, half, m.inst(Op.FMul, v4, oth_phi, c))
        m.stmt(Op.Branch, cont)
        m.label(cont)
        if op == "phi_swap":
            m.raw(Op.FAdd, v4, oth_next, val, m.const_fvec(0.0, 0.0, 0.0, 0.0))
        else:
            m.raw(Op.FAdd, v4, acc_next, val, m.const_fvec(0.0, 0.0, 0.0, 0.0))
            m.raw(Op.FAdd, v4, oth_next, oth_phi, m.const_fvec(0.0, 0.0, 0.0, 0.0))
        m.raw(Op.IAdd, it, i_next, i_phi, m.const_i(1))
        m.stmt(Op.Branch, head)
        m.label(merge)
        r = m.inst(Op.FMul, v4, m.inst(Op.FAdd, v4, acc_phi, oth_phi), m.const_fvec(0.25, 0.25, 0.25, 0.25))
    elif op in ("int_divmod", "uint_divmod", "shifts_bits", "int_cmp"):
        sv4 = m.t_vec(m.t_int(1), 4)
        uv4 = m.t_vec(m.t_int(0), 4)
        bv4 = m.t_vec(m.t_bool(), 4)

        def ints(x, scale, bias):    # (int)(x * scale) - bias: negative values and zeros included
            k = m.inst(Op.ConvertFToS, sv4, m.inst(Op.FMul, v4, x, m.const_fvec(scale, scale, scale, scale)))
            return m.inst(Op.ISub, sv4, k, m.inst(Op.ConvertFToS, sv4, m.const_fvec(bias, bias, bias, bias)))
        ia_, ib_ = ints(a, 4000.0, 900.0), ints(b, 9.0, 3.0)    # divisors in -3..6, zero and -1 among them
        if op == "int_divmod":
            x = m.inst(Op.IAdd, sv4, m.inst(Op.SDiv, sv4, ia_, ib_),
                       m.inst(Op.IAdd, sv4, m.inst(Op.IMul, sv4, m.inst(Op.SRem, sv4, ia_, ib_), m.inst(Op.ConvertFToS, sv4, m.const_fvec(7.0, 7.0, 7.0, 7.0))),
                              m.inst(Op.IMul, sv4, m.inst(Op.SMod, sv4, ia_, ib_), m.inst(Op.ConvertFToS, sv4, m.const_fvec(31.0, 31.0, 31.0, 31.0)))))
        elif op == "uint_divmod":
            ua, ub = m.inst(Op.Bitcast, uv4, ia_), m.inst(Op.Bitcast, uv4, m.inst(Op.BitwiseAnd, sv4, ib_, m.inst(Op.ConvertFToS, sv4, m.const_fvec(7.0, 7.0, 7.0, 7.0))))
            x = m.inst(Op.Bitcast, sv4, m.inst(Op.IAdd, uv4, m.inst(Op.UDiv, uv4, ua, ub), m.inst(Op.UMod, uv4, ua, ub)))
        elif op == "shifts_bits":
            sh = m.inst(Op.BitwiseAnd, sv4, ib_, m.inst(Op.ConvertFToS, sv4, m.const_fvec(63.0, 63.0, 63.0, 63.0)))
            t1 = m.inst(Op.ShiftRightArithmetic, sv4, ia_, sh)
            t2 = m.inst(Op.Bitcast, sv4, m.inst(Op.ShiftRightLogical, uv4, m.inst(Op.Bitcast, uv4, ia_), m.inst(Op.Bitcast, uv4, sh)))
            t3 = m.inst(Op.BitwiseXor, sv4, m.inst(Op.BitwiseOr, sv4, t1, ib_), m.inst(Op.Not, sv4, t2))
            x = m.inst(Op.IAdd, sv4, t3, m.inst(Op.SNegate, sv4, ia_))
        else:
            ua, ub = m.inst(Op.Bitcast, uv4, ia_), m.inst(Op.Bitcast, uv4, ib_)
            one = m.inst(Op.ConvertFToS, sv4, m.const_fvec(1.0, 1.0, 1.0, 1.0))
            zero = m.inst(Op.ConvertFToS, sv4, m.const_fvec(0.0, 0.0, 0.0, 0.0))
            x = zero
            for k, (cmp, l, rr) in enumerate(((Op.INotEqual, ia_, ib_), (Op.UGreaterThan, ua, ub), (Op.SGreaterThan, ia_, ib_),
                                              (Op.UGreaterThanEqual, ua, ub), (Op.SGreaterThanEqual, ia_, ib_),
                                              (Op.ULessThan, ua, ub), (Op.ULessThanEqual, ua, ub), (Op.SLessThanEqual, ia_, ib_))):
                bit = m.inst(Op.Select, sv4, m.inst(cmp, bv4, l, rr), one, zero)
                wgt = m.inst(Op.ConvertFToS, sv4, m.const_fvec(*([float(1 << k)] * 4)))
                x = m.inst(Op.IAdd, sv4, x, m.inst(Op.IMul, sv4, bit, wgt))
        x = m.inst(Op.BitwiseAnd, sv4, x, m.inst(Op.ConvertFToS, sv4, m.const_fvec(255.0, 255.0, 255.0, 255.0)))
        r = m.inst(Op.FMul, v4, m.inst(Op.ConvertSToF, v4, x), m.const_fvec(1 / 256.0, 1 / 256.0, 1 / 256.0, 1 / 256.0))
    elif op == "ucvt":
        uv4 = m.t_vec(m.t_int(0), 4)
        # a * 3e9 - 1e9: below zero, inside and above the uint range; back to float and scaled into [0, 1)
        x = m.inst(Op.FSub, v4, m.inst(Op.FMul, v4, a, m.const_fvec(3e9, 3e9, 3e9, 6e9)), m.const_fvec(1e9, 1e9, 1e9, 1e9))
        u = m.inst(Op.ConvertFToU, uv4, x)
        r = m.inst(Op.FMul, v4, m.inst(Op.ConvertUToF, v4, u), m.const_fvec(2.0 ** -32, 2.0 ** -32, 2.0 ** -32, 2.0 ** -32))
    elif op == "logic":
        bv4 = m.t_vec(m.t_bool(), 4)
        p_ = m.inst(Op.FOrdLessThan, bv4, a, b)
        q_ = m.inst(Op.FOrdLessThan, bv4, b, c)
        t = m.inst(Op.LogicalOr, bv4, m.inst(Op.LogicalAnd, bv4, p_, q_), m.inst(Op.LogicalNot, bv4, m.inst(Op.LogicalEqual, bv4, p_, q_)))
        t = m.inst(Op.LogicalNotEqual, bv4, t, m.inst(Op.FOrdLessThan, bv4, a, c))
        r = m.inst(Op.Select, v4, t, a, b)
    elif op == "isnan_inf":
        bv4 = m.t_vec(m.t_bool(), 4)
        zero = m.inst(Op.FSub, v4, a, a)
        nan = m.inst(Op.FDiv, v4, zero, zero)
        inf = m.inst(Op.FDiv, v4, b, zero)
        mixed = m.shuffle(v4, nan, inf, 0, 5, 2, 7)    # (nan, inf, nan, inf)
        probe = m.shuffle(v4, mixed, a, 0, 1, 6, 7)    # (nan, inf, a.z, a.w)
        r = m.inst(Op.Select, v4, m.inst(Op.IsNan, bv4, probe), a,
                   m.inst(Op.Select, v4, m.inst(Op.IsInf, bv4, probe), b, c))
    elif op == "switch_phi":
        # switch(int(a.x * 5)) { case 0: r = a; break; case 2: r = b; break; case 3: r = a + b; break; default: r = c; }
        # with the merge written as OpPhi, as a compiler's output has it
        it = m.t_int(1)
        sel = m.inst(Op.ConvertFToS, it, m.inst(Op.FMul, fl, ax, m.const_f(5.0)))
        c0, c2, c3, dflt, merge = (m.new_id() for _ in range(5))
        m.stmt(Op.SelectionMerge, merge, 0)
        m.stmt(Op.Switch, sel, dflt, 0, c0, 2, c2, 3, c3)
        m.label(c0)
        m.stmt(Op.Branch, merge)
        m.label(c2)
        m.stmt(Op.Branch, merge)
        m.label(c3)
        s3 = m.inst(Op.FAdd, v4, a, b)
        m.stmt(Op.Branch, merge)
        m.label(dflt)
        m.stmt(Op.Branch, merge)
        m.label(merge)
        r = m.new_id()
        m.raw(Op.Phi, v4, r, a, c0, b, c2, s3, c3, c, dflt)
    elif op == "consts_copy":
        # r = (true && a.x < b.x) ? copy(a) + null : (false || undef_bool ? c : b + undef_vec * 0)
        bt = m.t_bool()
        t_, f_ = m.const_bool(True), m.const_bool(False)
        cond = m.inst(Op.LogicalAnd, bt, t_, m.inst(Op.FOrdLessThan, bt, ax, bx))
        a2 = m.inst(Op.FAdd, v4, m.inst(Op.CopyObject, v4, a), m.const_null(v4))
        m.stmt(Op.Nop)
        ub = m.inst(Op.LogicalOr, bt, f_, m.inst(Op.Undef, bt))
        uv = m.inst(Op.FMul, v4, m.undef(v4), m.const_fvec(0.0, 0.0, 0.0, 0.0))
        alt = m.inst(Op.Select, v4, ub, c, m.inst(Op.FAdd, v4, b, uv))
        r = m.inst(Op.Select, v4, cond, a2, alt)
    elif op == "composite_insert":
        # v = a with .z replaced by b.x; M2 = M with column 1 replaced by v and element [2][3] by c.y; r = col sum
        v1 = m.inst(Op.CompositeInsert, v4, bx, a, 2)
        M2 = m.inst(Op.CompositeInsert, mat4, v1, M, 1)
        M3 = m.inst(Op.CompositeInsert, mat4, m.extract(fl, c, 1), M2, 2, 3)
        r = col_sum(M3)
    elif op == "vec_dynamic":
        # i = int(a.x * 6) (0..5: past the end included); r = insert(b, c[i], (i + 1) & 3) + splat(a[i])
        it = m.t_int(1)
        i = m.inst(Op.ConvertFToS, it, m.inst(Op.FMul, fl, ax, m.const_f(6.0)))
        e = m.inst(Op.VectorExtractDynamic, fl, c, i)
        j = m.inst(Op.BitwiseAnd, it, m.inst(Op.IAdd, it, i, m.const_i(1)), m.const_i(3))
        ins = m.inst(Op.VectorInsertDynamic, v4, b, e, j)
        ins2 = m.inst(Op.VectorInsertDynamic, v4, ins, ax, m.inst(Op.IAdd, it, i, m.const_i(2)))    # may be past the end
        r = m.inst(Op.FAdd, v4, ins2, splat(m.inst(Op.VectorExtractDynamic, fl, a, i)))
    elif op == "frem_fmod":
        # x = b * 7.3 (both signs), y = c (> 0) and -c: remainder with the sign of x, modulo with the sign of y
        x = m.inst(Op.FMul, v4, b, m.const_fvec(7.3, 7.3, 7.3, 7.3))
        rem = m.inst(Op.FRem, v4, x, c)
        mod = m.inst(Op.FMod, v4, x, m.inst(Op.FNegate, v4, c))
        r = m.inst(Op.FAdd, v4, m.inst(Op.FMul, v4, rem, m.const_fvec(0.5, 0.5, 0.5, 0.5)),
                   m.inst(Op.FMul, v4, mod, m.const_fvec(-0.5, -0.5, -0.5, -0.5)))
    elif op == "any_all":
        bt = m.t_bool()
        bv4 = m.t_vec(bt, 4)
        any_ = m.inst(Op.Any, bt, m.inst(Op.FOrdLessThan, bv4, a, b))
        all_ = m.inst(Op.All, bt, m.inst(Op.FOrdLessThanEqual, bv4, c, b))
        r = m.inst(Op.FAdd, v4, m.inst(Op.FMul, v4, m.inst(Op.Select, v4, any_, b, c), m.const_fvec(0.5, 0.5, 0.5, 0.5)),
                   m.inst(Op.FMul, v4, m.inst(Op.Select, v4, all_, c, a), m.const_fvec(0.25, 0.25, 0.25, 0.25)))
    elif op == "bit_ops":
        sv4 = m.t_vec(m.t_int(1), 4)
        uv4 = m.t_vec(m.t_int(0), 4)

        def ints(x, scale, bias):
            k = m.inst(Op.ConvertFToS, sv4, m.inst(Op.FMul, v4, x, m.const_fvec(scale, scale, scale, scale)))
            return m.inst(Op.ISub, sv4, k, m.inst(Op.ConvertFToS, sv4, m.const_fvec(bias, bias, bias, bias)))

        def ci(val):
            return m.inst(Op.ConvertFToS, sv4, m.const_fvec(*([float(val)] * 4)))
        big, small = ints(b, 4000.0, 900.0), ints(c, 9.0, 3.0)    # small: -3..5, zero and -1 among them
        x = m.inst(Op.BitCount, sv4, big)
        rv = m.inst(Op.Bitcast, sv4, m.inst(Op.ShiftRightLogical, uv4, m.inst(Op.Bitcast, uv4, m.inst(Op.BitReverse, sv4, big)),
                                            m.inst(Op.Bitcast, uv4, ci(24))))
        x = m.inst(Op.IAdd, sv4, x, rv)
        for code, arg, wgt in ((GLSL.FindILsb, big, 3), (GLSL.FindILsb, small, 5), (GLSL.FindSMsb, big, 7),
                               (GLSL.FindSMsb, small, 11), (GLSL.FindUMsb, big, 13), (GLSL.FindUMsb, small, 17)):
            x = m.inst(Op.IAdd, sv4, x, m.inst(Op.IMul, sv4, m.ext(sv4, code, arg), ci(wgt)))
        x = m.inst(Op.BitwiseAnd, sv4, x, ci(255))
        r = m.inst(Op.FMul, v4, m.inst(Op.ConvertSToF, v4, x), m.const_fvec(1 / 256.0, 1 / 256.0, 1 / 256.0, 1 / 256.0))
    elif op == "nminmax":
        zero = m.inst(Op.FSub, v4, a, a)
        nan = m.inst(Op.FDiv, v4, zero, zero)
        probe = m.shuffle(v4, nan, b, 0, 5, 2, 7)    # (nan, b.y, nan, b.w)
        other = m.shuffle(v4, c, nan, 0, 1, 6, 3)    # (c.x, c.y, nan, c.w): one lane has NaN on both sides
        lo = m.inst(Op.FMul, v4, c, m.const_fvec(0.5, 0.5, 0.5, 0.5))
        t1 = m.ext(v4, GLSL.NMin, probe, c)
        t2 = m.ext(v4, GLSL.NMax, c, probe)
        t3 = m.ext(v4, GLSL.NClamp, probe, lo, c)
        t4 = m.ext(v4, GLSL.NMin, probe, other)
        # NaN survives only where both operands were NaN (lane 2 of t4): replaced through IsNan so the colour is defined
        bv4 = m.t_vec(m.t_bool(), 4)
        t4 = m.inst(Op.Select, v4, m.inst(Op.IsNan, bv4, t4), a, t4)
        r = m.inst(Op.FAdd, v4, m.inst(Op.FAdd, v4, m.inst(Op.FMul, v4, t1, m.const_fvec(0.25, 0.25, 0.25, 0.25)),
                                       m.inst(Op.FMul, v4, t2, m.const_fvec(0.25, 0.25, 0.25, 0.25))),
                   m.inst(Op.FAdd, v4, m.inst(Op.FMul, v4, t3, m.const_fvec(0.25, 0.25, 0.25, 0.25)),
                          m.inst(Op.FMul, v4, t4, m.const_fvec(0.125, 0.125, 0.125, 0.125))))
    elif op == "exp_log":
        t = m.inst(Op.FAdd, v4, m.ext(v4, GLSL.FAbs, b), m.const_fvec(0.5, 0.5, 0.5, 0.5))
        bh = m.inst(Op.FMul, v4, b, m.const_fvec(0.5, 0.5, 0.5, 0.5))
        terms = ((m.ext(v4, GLSL.Exp, bh), 0.15), (m.ext(v4, GLSL.Exp2, b), 0.1), (m.ext(v4, GLSL.Log, t), 0.08),
                 (m.ext(v4, GLSL.Log2, t), 0.05))
        r = m.const_fvec(0.2, 0.2, 0.2, 0.2)
        for val, wgt in terms:
            r = m.inst(Op.FAdd, v4, r, m.inst(Op.FMul, v4, val, m.const_fvec(wgt, wgt, wgt, wgt)))
    elif op == "tan_hyp":
        t = m.ext(v4, GLSL.Fract, b)
        terms = ((m.ext(v4, GLSL.Tan, t), 0.15), (m.ext(v4, GLSL.Sinh, t), 0.15), (m.ext(v4, GLSL.Cosh, t), 0.15),
                 (m.ext(v4, GLSL.Tanh, m.inst(Op.FMul, v4, b, m.const_fvec(3.0, 3.0, 3.0, 3.0))), 0.1))
        r = m.const_fvec(0.1, 0.1, 0.1, 0.1)
        for val, wgt in terms:
            r = m.inst(Op.FAdd, v4, r, m.inst(Op.FMul, v4, val, m.const_fvec(wgt, wgt, wgt, wgt)))
    elif op == "atan_asin":
        t = m.inst(Op.FSub, v4, m.inst(Op.FMul, v4, m.ext(v4, GLSL.Fract, b), m.const_fvec(1.8, 1.8, 1.8, 1.8)),
                   m.const_fvec(0.9, 0.9, 0.9, 0.9))
        x2 = m.inst(Op.FSub, v4, c, m.const_fvec(0.5, 0.5, 0.5, 0.5))
        terms = ((m.ext(v4, GLSL.Atan, m.inst(Op.FMul, v4, b, m.const_fvec(3.0, 3.0, 3.0, 3.0))), 0.1),
                 (m.ext(v4, GLSL.Atan2, b, x2), 0.05), (m.ext(v4, GLSL.Asin, t), 0.1), (m.ext(v4, GLSL.Acos, t), 0.08))
        r = m.const_fvec(0.4, 0.4, 0.4, 0.4)
        for val, wgt in terms:
            r = m.inst(Op.FAdd, v4, r, m.inst(Op.FMul, v4, val, m.const_fvec(wgt, wgt, wgt, wgt)))
    elif op == "bitfield":
        it = m.t_int(1)
        sv4 = m.t_vec(it, 4)

        def ints(x, scale, bias):
            k = m.inst(Op.ConvertFToS, sv4, m.inst(Op.FMul, v4, x, m.const_fvec(scale, scale, scale, scale)))
            return m.inst(Op.ISub, sv4, k, m.inst(Op.ConvertFToS, sv4, m.const_fvec(bias, bias, bias, bias)))

        def ci(val):
            return m.inst(Op.ConvertFToS, sv4, m.const_fvec(*([float(val)] * 4)))
        big, ins = ints(b, 400000.0, 90000.0), ints(c, 4000.0, 900.0)
        # offset 0..15 from a.x, count 0..16 from c.x: offset + count <= 31; scalars, as the instruction wants them
        off = m.inst(Op.BitwiseAnd, it, m.inst(Op.ConvertFToS, it, m.inst(Op.FMul, fl, m.ext(fl, GLSL.FAbs, ax), m.const_f(97.0))),
                     m.const_i(15))
        cnt = m.inst(Op.ConvertFToS, it, m.inst(Op.FMul, fl, m.extract(fl, c, 0), m.const_f(16.9)))
        x = m.inst(Op.BitFieldSExtract, sv4, big, off, cnt)
        x = m.inst(Op.IAdd, sv4, x, m.inst(Op.IMul, sv4, m.inst(Op.BitFieldUExtract, sv4, big, off, cnt), ci(3)))
        x = m.inst(Op.IAdd, sv4, x, m.inst(Op.ShiftRightArithmetic, sv4, m.inst(Op.BitFieldInsert, sv4, big, ins, off, cnt), ci(5)))
        x = m.inst(Op.BitwiseAnd, sv4, x, ci(255))
        r = m.inst(Op.FMul, v4, m.inst(Op.ConvertSToF, v4, x), m.const_fvec(1 / 256.0, 1 / 256.0, 1 / 256.0, 1 / 256.0))
    elif op == "determinant":
        mat3 = m.t_mat(3)
        c3 = m.shuffle(v3, c, c, 0, 1, 2)
        M3 = m.construct(mat3, a3, b3, c3)
        d4m, d4n, d3 = m.ext(fl, GLSL.Determinant, M), m.ext(fl, GLSL.Determinant, N), m.ext(fl, GLSL.Determinant, M3)
        r = m.inst(Op.FAdd, v4, m.const_fvec(0.5, 0.5, 0.5, 0.5),
                   m.inst(Op.FMul, v4, m.construct(v4, d4m, d4n, d3, m.inst(Op.FAdd, fl, d4m, d3)), m.const_fvec(8.0, 8.0, 0.4, 0.4)))
    elif op == "funord":
        # the six unordered comparisons of (NaN, b.y, b.z, b.w) with c (c == b on every third vertex), as bit weights
        bv4 = m.t_vec(m.t_bool(), 4)
        zero = m.inst(Op.FSub, v4, a, a)
        nan = m.inst(Op.FDiv, v4, zero, zero)
        probe = m.shuffle(v4, nan, b, 0, 5, 6, 7)
        r = m.const_fvec(0.0, 0.0, 0.0, 0.0)
        for k, cmp in enumerate((Op.FUnordEqual, Op.FUnordNotEqual, Op.FUnordLessThan, Op.FUnordGreaterThan,
                                 Op.FUnordLessThanEqual, Op.FUnordGreaterThanEqual)):
            wgt = 2.0 ** -(k + 1)
            r = m.inst(Op.FAdd, v4, r, m.inst(Op.Select, v4, m.inst(cmp, bv4, probe, c), m.const_fvec(wgt, wgt, wgt, wgt),
                                              m.const_fvec(0.0, 0.0, 0.0, 0.0)))
    elif op == "inv_hyp":
        t = m.ext(v4, GLSL.Fract, b)    # [0, 1)
        terms = ((m.ext(v4, GLSL.Asinh, m.inst(Op.FMul, v4, b, m.const_fvec(2.0, 2.0, 2.0, 2.0))), 0.1),
                 (m.ext(v4, GLSL.Acosh, m.inst(Op.FAdd, v4, t, m.const_fvec(1.25, 1.25, 1.25, 1.25))), 0.2),
                 (m.ext(v4, GLSL.Atanh, m.inst(Op.FMul, v4, t, m.const_fvec(0.9, 0.9, 0.9, 0.9))), 0.15))
        r = m.const_fvec(0.35, 0.35, 0.35, 0.35)
        for val, wgt in terms:
            r = m.inst(Op.FAdd, v4, r, m.inst(Op.FMul, v4, val, m.const_fvec(wgt, wgt, wgt, wgt)))
    elif op == "fabs":
        r = m.ext(v4, GLSL.FAbs, a)
    elif op == "floor":
        r = m.ext(v4, GLSL.Floor, a)
    elif op == "fract":
        r = m.ext(v4, GLSL.Fract, a)
    else:
        raise ValueError(op)
    m.store(m.access(SC.Output, v4, gl, m.const_i(0)), a)
    m.store(out, r)
    return _finish(m, VERTEX, f, [ia, ib, ic, out, gl] + extra_iface)
