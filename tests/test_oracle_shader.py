"""Known-answer tests that pin oracle/spirv_cpu.cpp (the CPU restatement of the reference's SPIR-V
front end + JIT wrappers, spirv_compile.cpp) op by op against an independent numpy float32
re-computation written from the reference's semantics.

The reference's LLVM-6 JIT cannot be built here, so this is what pins the shader stage
("parity unpinned" against the reference's own binary; pinned against its documented arithmetic).
"""
import ctypes as C

import numpy as np
import pytest

from harness import abi, shaders

f32 = np.float32


def _dot(a, b, n):
    acc = f32(a[0] * b[0])
    for i in range(1, n):
        acc = f32(acc + f32(a[i] * b[i]))
    return acc


def _mxv(Mc, v):  # Mc[col][row]; Float4x4TimesVec4: out[r] = 0; out[r] += m[c*4+r]*v[c]
    out = []
    for r in range(4):
        acc = f32(0.0)
        for c in range(4):
            acc = f32(acc + f32(Mc[c][r] * v[c]))
        out.append(acc)
    return np.array(out, dtype=f32)


def _vxm(Mc, v):  # Vec4TimesFloat4x4: out[r] += m[r*4+c]*v[c]
    out = []
    for r in range(4):
        acc = f32(0.0)
        for c in range(4):
            acc = f32(acc + f32(Mc[r][c] * v[c]))
        out.append(acc)
    return np.array(out, dtype=f32)


def _mxm(A, B):  # Float4x4TimesFloat4x4(a, b): out[x][y] = ((b[x][0]*a[0][y] + b[x][1]*a[1][y]) + ..)
    out = np.zeros((4, 4), dtype=f32)
    for x in range(4):
        for y in range(4):
            acc = f32(B[x][0] * A[0][y])
            for k in range(1, 4):
                acc = f32(acc + f32(B[x][k] * A[k][y]))
            out[x][y] = acc
    return out


def expected(op, a, b, c, M, N):
    one = f32(1.0)
    if op == "fadd":
        return a + b
    if op == "fsub":
        return a - b
    if op == "fmul":
        return a * b
    if op == "fdiv":
        return a / b
    if op == "fneg":
        return f32(-0.0) - a
    if op == "vts":
        return a * b[0]
    if op == "dot4":
        return np.full(4, _dot(a, b, 4), dtype=f32)
    if op == "dot3":
        return np.full(4, _dot(a, b, 3), dtype=f32)
    if op == "fmin":
        return np.where(a < b, a, b)
    if op == "fmax":
        return np.where(a > b, a, b)
    if op == "fclamp":  # val=a, lower=b, upper=c
        up = np.where(a < c, a, c)
        return np.where(up > b, up, b)
    if op == "fmix":  # x=a, y=b, a=c: (1-c)*x + c*y
        return (one - c) * a + c * b
    if op == "sqrt":
        return np.full(4, np.sqrt(a[0]), dtype=f32)
    if op == "invsqrt":
        return np.full(4, one / np.sqrt(a[0]), dtype=f32)
    if op == "normalize3":
        inv = one / np.sqrt(_dot(a, a, 3))
        return np.array([a[0] * inv, a[1] * inv, a[2] * inv, 0], dtype=f32)
    if op == "length3":
        return np.full(4, np.sqrt(_dot(a, a, 3)), dtype=f32)
    if op == "reflect3":
        d2 = f32(_dot(a, b, 3) * f32(2.0))
        return np.array([a[i] - f32(d2 * b[i]) for i in range(3)] + [0], dtype=f32)
    if op == "cross3":  # the reference returns operand 0 (spirv_compile.cpp:1661)
        return np.array([a[0], a[1], a[2], 0], dtype=f32)
    if op == "shuffle":
        return np.array([a[3], b[0], a[1], b[2]], dtype=f32)
    if op == "mxv":
        return _mxv(M, a)
    if op == "vxm":
        return _vxm(M, a)
    if op == "mxm":
        P = _mxm(M, N)
        r = P[0].copy()
        for i in (1, 2, 3):
            r = r + P[i]
        return r
    if op == "transpose":
        return np.array([M[r][1] for r in range(4)], dtype=f32)
    if op == "mxs":
        return M[2] * a[0]
    if op == "minverse":  # the reference calls Float4x4Transpose (spirv_compile.cpp:1721)
        return np.array([M[r][3] for r in range(4)], dtype=f32)
    if op == "sin":
        return np.full(4, np.sin(a[0]), dtype=f32)
    if op == "cos":
        return np.full(4, np.cos(a[0]), dtype=f32)
    if op == "pow":
        return np.power(a, b).astype(f32)
    raise ValueError(op)


def unit_inputs(seed=0, n=64):
    rng = np.random.default_rng(seed)
    verts = rng.uniform(0.1, 3.0, size=(n, 12)).astype(f32)
    verts[:, 0:4] *= rng.choice([-1.0, 1.0], size=(n, 4)).astype(f32)
    verts[:, 4:8] *= rng.choice([-1.0, 1.0], size=(n, 4)).astype(f32)
    ubo = rng.uniform(-2.0, 2.0, size=32).astype(f32)
    return verts, ubo


def unit_state(be, op, verts, ubo):
    """A draw state good enough for vor_run_vertex: three vec4 attributes from one interleaved VB."""
    mod = be.CompileFunction(shaders.vs_unit(op))
    pl = abi.Pipeline()
    for loc in range(3):
        pl.vattrs[loc].format, pl.vattrs[loc].stride = abi.FMT_R32G32B32A32_SFLOAT, 48
        pl.vattrs[loc].offset, pl.vattrs[loc].vb = 16 * loc, 0
    st = abi.DrawState()
    st.vbs[0].buffer = abi.make_buffer(verts)
    st.pipeline = C.pointer(pl)
    arr = (abi.Binding * 1)()
    arr[0].set, arr[0].binding, arr[0].type = 0, 0, abi.DESC_UNIFORM_BUFFER
    arr[0].buffer = abi.make_buffer(ubo)
    st.bindings = C.cast(arr, C.POINTER(abi.Binding))
    st.num_bindings = 1
    return mod, st, (pl, arr)


@pytest.mark.parametrize("op", shaders.UNIT_OPS)
def test_unit_op_matches_numpy(vor, op):
    verts, ubo = unit_inputs()
    if op in ("sqrt", "invsqrt", "pow"):
        verts[:, 0:4] = np.abs(verts[:, 0:4])
    mod, st, keep = unit_state(vor, op, verts, ubo)
    entry = vor.GetFuncPointer(mod, "main")
    run = vor.lib.vor_run_vertex
    run.argtypes = [C.POINTER(abi.DrawState), C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    M = ubo[:16].reshape(4, 4)  # M[col][row]
    N = ubo[16:].reshape(4, 4)
    out = (C.c_float * 44)()
    for i in range(verts.shape[0]):
        assert run(C.byref(st), entry, i, out) == 0
        got = np.array(out[4:8], dtype=f32)
        pos = np.array(out[0:4], dtype=f32)
        a, b, c = verts[i, 0:4], verts[i, 4:8], verts[i, 8:12]
        assert np.array_equal(pos.view(np.uint32), a.view(np.uint32))
        with np.errstate(all="ignore"):
            exp = np.asarray(expected(op, a, b, c, M, N), dtype=f32)
        if op in ("sin", "cos", "pow"):  # libm vs numpy: not bit-reproducible by construction
            assert np.allclose(got, exp, rtol=1e-5, atol=1e-6)
        else:
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (op, i, got, exp)
    vor.DestroyFunction(mod)


def test_dynamic_indices_into_local_arrays(vor):
    """OpAccessChain with run-time indices into function-local arrays and vectors (a GEP in the reference,
    spirv_compile.cpp:1301-1318): loads and stores through them, against a numpy restatement"""
    verts, ubo = unit_inputs(11)
    mod, st, keep = unit_state(vor, "dynidx", verts, ubo)
    entry = vor.GetFuncPointer(mod, "main")
    run = vor.lib.vor_run_vertex
    run.argtypes = [C.POINTER(abi.DrawState), C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    out = (C.c_float * 44)()
    for n in range(verts.shape[0]):
        assert run(C.byref(st), entry, n, out) == 0
        got = np.array(out[4:8], dtype=f32)
        a, b, c = verts[n, 0:4], verts[n, 4:8], verts[n, 8:12]
        i, j = n & 3, (n * 3) & 3
        arr = [a, b, c, (a + b).astype(f32)]
        fa = [a[0], b[1], c[2], a[3]]
        fa[i] = c[0]
        exp = (arr[i] + np.array([fa[1], fa[j], b[i], fa[1]], dtype=f32)).astype(f32)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (n, got, exp)
    vor.DestroyFunction(mod)


_UBO_FOR_EXT = None    # the uniform block of the running test (the matrix operand of composite_insert)


def _expected_ext(op, a, b, c):
    f32 = np.float32
    if op == "select":
        return np.where(a < b, a, c)
    if op == "fge":
        return np.where(a >= b, a, c)
    if op == "feq":
        return np.where(a == b, a, c)
    if op == "fne":
        return np.where(a != b, a, c)
    if op == "isub_bitcast":
        return (a.view(np.int32) - b.view(np.int32)).view(f32)
    if op == "ftos":
        return np.trunc(a).astype(np.int32).astype(f32)
    if op == "fabs":
        return np.abs(a)
    if op == "floor":
        return np.floor(a)
    if op == "fract":
        return a - np.floor(a)
    if op in ("roundeven", "trunc", "ceil"):
        x = (a * np.array([3.5, 2.5, 7.25, 0.5], f32)).astype(f32)
        fn = {"roundeven": np.rint, "trunc": np.trunc, "ceil": np.ceil}[op]
        return (fn(x).astype(f32) * f32(0.125)).astype(f32)
    if op == "fsign":
        return np.where(a > 0, f32(1), np.where(a < 0, f32(-1), f32(0))).astype(f32)
    if op == "radians":
        return (a * f32(0.017453292519943295)).astype(f32)
    if op == "degrees":
        return ((a * f32(57.29577951308232)).astype(f32) * f32(0.125)).astype(f32) * f32(0.125)
    if op == "step":
        return np.where(b < a, f32(0), f32(1)).astype(f32)
    if op == "smoothstep":
        with np.errstate(all="ignore"):
            q = ((c - a).astype(f32) / (b - a).astype(f32)).astype(f32)
        u = np.where(q < 1, q, f32(1)).astype(f32)
        t = np.where(u > 0, u, f32(0)).astype(f32)
        return ((t * t).astype(f32) * (f32(3) - (f32(2) * t).astype(f32)).astype(f32)).astype(f32)
    if op == "fma":
        return np.array([np.float64(x) * np.float64(y) + np.float64(z) for x, y, z in zip(a, b, c)]).astype(f32)
    def dot3(x, y):    # CreateDot order: ((x0*y0 + x1*y1) + x2*y2)
        return f32(f32(f32(x[0] * y[0]) + f32(x[1] * y[1])) + f32(x[2] * y[2]))
    if op == "distance3":
        d = (a[:3] - b[:3]).astype(f32)
        return np.full(4, np.sqrt(dot3(d, d)), f32)
    if op == "faceforward3":
        nref = (c[:3] - a[:3]).astype(f32)
        dd = dot3(nref, b[:3])
        r = a[:3] if dd < 0 else (f32(-0.0) - a[:3]).astype(f32)
        return np.array([r[0], r[1], r[2], 0.0], f32)
    if op == "refract3":
        def normalize(v):
            inv = f32(f32(1.0) / np.sqrt(dot3(v, v)))
            return (v * inv).astype(f32)
        I, N, eta = normalize(a[:3]), normalize(b[:3]), c[0]
        d = dot3(N, I)
        kk = f32(f32(1) - f32(f32(eta * eta) * f32(f32(1) - f32(d * d))))
        if kk < 0:
            return np.zeros(4, f32)
        t = f32(f32(eta * d) + np.sqrt(kk))
        r = ((eta * I).astype(f32) - (t * N).astype(f32)).astype(f32)
        return np.array([r[0], r[1], r[2], 0.0], f32)
    if op in ("int_minmax", "uint_minmax", "int_abs_sign"):
        def ints(x):
            return np.trunc((x * f32(64)).astype(f32)).astype(np.int64) - 24
        ia, ib, ic = ints(a), ints(b), ints(c)
        if op == "int_abs_sign":
            x = np.abs(ia) + np.sign(ib)
        else:
            if op == "uint_minmax":
                ia, ib, ic = ia & 0xffffffff, ib & 0xffffffff, ic & 0xffffffff
            lo, hi = np.minimum(ia, ib), np.maximum(ia, ib)
            x = lo + hi + np.minimum(np.maximum(ic, lo), hi)
        return ((x & 255).astype(f32) * f32(1 / 256.0)).astype(f32)
    if op in ("int_divmod", "uint_divmod", "shifts_bits", "int_cmp"):
        def ints(x, scale, bias):
            return np.trunc((x * f32(scale)).astype(f32)).astype(np.int64) - bias
        ia, ib = ints(a, 4000.0, 900), ints(b, 9.0, 3)

        def wrap(v):    # int64 -> two's complement int32
            return ((v + 2 ** 31) % 2 ** 32) - 2 ** 31

        def tdiv(x, y):    # C division truncating toward zero
            return np.where(y == 0, 0, np.sign(x) * np.sign(np.where(y == 0, 1, y)) * (np.abs(x) // np.abs(np.where(y == 0, 1, y))))
        if op == "int_divmod":
            q = np.where(ib == 0, 0, np.where(ib == -1, wrap(-ia), tdiv(ia, ib)))
            rem = np.where((ib == 0) | (ib == -1), 0, ia - tdiv(ia, ib) * ib)
            mod = np.where((rem != 0) & ((rem ^ ib) < 0), rem + ib, rem)
            x = wrap(q + wrap(rem * 7) + wrap(mod * 31))
        elif op == "uint_divmod":
            ua, ub = ia & 0xffffffff, (ib & 7) & 0xffffffff
            x = np.where(ub == 0, 0, ua // np.where(ub == 0, 1, ub)) + np.where(ub == 0, 0, ua % np.where(ub == 0, 1, ub))
        elif op == "shifts_bits":
            sh = (ib & 63) & 31
            t1 = ia >> sh
            t2 = wrap((ia & 0xffffffff) >> sh)
            t3 = (t1 | ib) ^ wrap(~t2)
            x = wrap(t3 + wrap(-ia))
        else:
            ua, ub = ia & 0xffffffff, ib & 0xffffffff
            bits = [ia != ib, ua > ub, ia > ib, ua >= ub, ia >= ib, ua < ub, ua <= ub, ia <= ib]
            x = sum(bt.astype(np.int64) << k for k, bt in enumerate(bits))
        return ((x & 255).astype(f32) * f32(1 / 256.0)).astype(f32)
    if op == "ucvt":
        x = ((a * np.array([3e9, 3e9, 3e9, 6e9], f32)).astype(f32) - f32(1e9)).astype(f32)
        u = np.where((x > -1) & (x < 4294967296.0), np.trunc(np.maximum(x, 0)), 0).astype(np.uint32)
        return (u.astype(f32) * f32(2.0 ** -32)).astype(f32)
    if op == "logic":
        p_, q_ = a < b, b < c
        t = (p_ & q_) | ~(p_ == q_)
        t = t != (a < c)
        return np.where(t, a, b)
    if op == "isnan_inf":
        return np.array([a[0], b[1], c[2], c[3]], f32)
    if op == "switch_phi":
        sel = int(np.trunc(f32(a[0] * f32(5.0))))
        return {0: a, 2: b, 3: (a + b).astype(f32)}.get(sel, c)
    if op == "consts_copy":
        return a if a[0] < b[0] else b
    if op == "composite_insert":
        M = np.array(_UBO_FOR_EXT[:16], f32).reshape(4, 4).copy()    # M[col][row]
        v1 = a.copy()
        v1[2] = b[0]
        M[1] = v1
        M[2][3] = c[1]
        r = M[0].copy()
        for k in (1, 2, 3):
            r = (r + M[k]).astype(f32)
        return r
    if op == "vec_dynamic":
        i = int(np.trunc(f32(a[0] * f32(6.0))))
        e = c[i] if 0 <= i < 4 else c[0]
        ins = b.copy()
        ins[(i + 1) & 3] = e
        if 0 <= i + 2 < 4:
            ins[i + 2] = a[0]
        return (ins + (a[i] if 0 <= i < 4 else a[0])).astype(f32)
    if op == "frem_fmod":
        x = (b * f32(7.3)).astype(f32)
        rem = (x - (c * np.trunc((x / c).astype(f32)).astype(f32)).astype(f32)).astype(f32)
        nc = (-c).astype(f32)
        mod = (x - (nc * np.floor((x / nc).astype(f32)).astype(f32)).astype(f32)).astype(f32)
        return ((rem * f32(0.5)).astype(f32) + (mod * f32(-0.5)).astype(f32)).astype(f32)
    if op == "any_all":
        s1 = b if np.any(a < b) else c
        s2 = c if np.all(c <= b) else a
        return ((s1 * f32(0.5)).astype(f32) + (s2 * f32(0.25)).astype(f32)).astype(f32)
    if op == "bit_ops":
        def ints(x, scale, bias):
            return (np.trunc((x * f32(scale)).astype(f32)).astype(np.int64) - bias).astype(np.int32)
        big, small = ints(b, 4000.0, 900), ints(c, 9.0, 3)
        def lsb(v):
            u = int(v) & 0xffffffff
            return -1 if u == 0 else (u & -u).bit_length() - 1
        def umsb(v):
            return (int(v) & 0xffffffff).bit_length() - 1
        def smsb(v):
            v = int(v)
            return umsb(~v if v < 0 else v)
        out = []
        for k in range(4):
            ub = int(big[k]) & 0xffffffff
            x = bin(ub).count("1") + (int(f"{ub:032b}"[::-1], 2) >> 24)
            x += 3 * lsb(big[k]) + 5 * lsb(small[k]) + 7 * smsb(big[k]) + 11 * smsb(small[k])
            x += 13 * umsb(big[k]) + 17 * umsb(small[k])
            out.append(f32(x & 255) * f32(1 / 256.0))
        return np.array(out, f32)
    if op == "nminmax":
        nan = f32(np.nan)
        probe = np.array([nan, b[1], nan, b[3]], f32)
        other = np.array([c[0], c[1], nan, c[3]], f32)
        def nsel(is_min, x, y):
            if x != x:
                return y
            if y != y:
                return x
            return y if ((y < x) if is_min else (y > x)) else x
        lo = (c * f32(0.5)).astype(f32)
        t1 = np.array([nsel(True, probe[k], c[k]) for k in range(4)], f32)
        t2 = np.array([nsel(False, c[k], probe[k]) for k in range(4)], f32)
        t3 = np.array([nsel(True, nsel(False, probe[k], lo[k]), c[k]) for k in range(4)], f32)
        t4 = np.array([nsel(True, probe[k], other[k]) for k in range(4)], f32)
        t4 = np.where(np.isnan(t4), a, t4).astype(f32)
        q = f32(0.25)
        return (((t1 * q).astype(f32) + (t2 * q).astype(f32)).astype(f32)
                + ((t3 * q).astype(f32) + (t4 * f32(0.125)).astype(f32)).astype(f32)).astype(f32)
    if op == "bitfield":
        def ints(x, scale, bias):
            return (np.trunc((x * f32(scale)).astype(f32)).astype(np.int64) - bias).astype(np.int32)
        big, ins = ints(b, 400000.0, 90000), ints(c, 4000.0, 900)
        off = int(np.trunc(f32(np.abs(a[0]) * f32(97.0)))) & 15
        cnt = int(np.trunc(f32(c[0] * f32(16.9))))
        mask = (1 << cnt) - 1
        out = []
        for k in range(4):
            ub, ui = int(big[k]) & 0xffffffff, int(ins[k]) & 0xffffffff
            ux = (ub >> off) & mask
            sx = ux - (1 << cnt) if cnt and (ux >> (cnt - 1)) & 1 else ux
            inserted = (ub & ~(mask << off) & 0xffffffff) | ((ui << off) & (mask << off) & 0xffffffff)
            if inserted & 0x80000000:
                inserted -= 1 << 32
            x = sx + 3 * ux + (inserted >> 5)
            out.append(f32(x & 255) * f32(1 / 256.0))
        return np.array(out, f32)
    if op == "determinant":
        M = np.array(_UBO_FOR_EXT[:16], f32).reshape(4, 4)     # [col][row]
        N = np.array(_UBO_FOR_EXT[16:32], f32).reshape(4, 4)
        M3 = np.array([a[:3], b[:3], c[:3]], f32)              # columns a, b, c
        def det3(e, rw, cl):
            m0 = f32(f32(e(rw[1], cl[1]) * e(rw[2], cl[2])) - f32(e(rw[2], cl[1]) * e(rw[1], cl[2])))
            m1 = f32(f32(e(rw[1], cl[0]) * e(rw[2], cl[2])) - f32(e(rw[2], cl[0]) * e(rw[1], cl[2])))
            m2 = f32(f32(e(rw[1], cl[0]) * e(rw[2], cl[1])) - f32(e(rw[2], cl[0]) * e(rw[1], cl[1])))
            d = f32(e(rw[0], cl[0]) * m0)
            d = f32(d - f32(e(rw[0], cl[1]) * m1))
            return f32(d + f32(e(rw[0], cl[2]) * m2))
        def det4(mat):
            e = lambda r_, c_: mat[c_][r_]
            d = None
            for j in range(4):
                cl = [c_ for c_ in range(4) if c_ != j]
                t = f32(e(0, j) * det3(e, (1, 2, 3), cl))
                d = t if j == 0 else (f32(d - t) if j & 1 else f32(d + t))
            return d
        d4m, d4n = det4(M), det4(N)
        d3 = det3(lambda r_, c_: M3[c_][r_], (0, 1, 2), (0, 1, 2))
        v = np.array([d4m, d4n, d3, f32(d4m + d3)], f32)
        return (f32(0.5) + (v * np.array([8.0, 8.0, 0.4, 0.4], f32)).astype(f32)).astype(f32)
    if op == "funord":
        probe = np.array([np.nan, b[1], b[2], b[3]], f32)
        r = np.zeros(4, f32)
        with np.errstate(invalid="ignore"):
            tests = (~((probe < c) | (probe > c)), ~(probe == c), ~(probe >= c), ~(probe <= c), ~(probe > c), ~(probe < c))
        for k, t in enumerate(tests):
            r = (r + np.where(t, f32(2.0 ** -(k + 1)), f32(0))).astype(f32)
        return r
    if op in ("exp_log", "tan_hyp", "atan_asin", "inv_hyp"):    # approximate: compared with allclose by the caller
        if op == "inv_hyp":
            t = (b - np.floor(b)).astype(f32)
            terms = ((np.arcsinh((b * f32(2.0)).astype(f32)), 0.1), (np.arccosh((t + f32(1.25)).astype(f32)), 0.2),
                     (np.arctanh((t * f32(0.9)).astype(f32)), 0.15))
            r = np.full(4, 0.35, f32)
        elif op == "exp_log":
            t = (np.abs(b) + f32(0.5)).astype(f32)
            terms = ((np.exp((b * f32(0.5)).astype(f32)), 0.15), (np.exp2(b), 0.1), (np.log(t), 0.08), (np.log2(t), 0.05))
            r = np.full(4, 0.2, f32)
        elif op == "tan_hyp":
            t = (b - np.floor(b)).astype(f32)
            terms = ((np.tan(t), 0.15), (np.sinh(t), 0.15), (np.cosh(t), 0.15), (np.tanh((b * f32(3.0)).astype(f32)), 0.1))
            r = np.full(4, 0.1, f32)
        else:
            t = (((b - np.floor(b)).astype(f32) * f32(1.8)).astype(f32) - f32(0.9)).astype(f32)
            x2 = (c - f32(0.5)).astype(f32)
            terms = ((np.arctan((b * f32(3.0)).astype(f32)), 0.1), (np.arctan2(b, x2), 0.05), (np.arcsin(t), 0.1),
                     (np.arccos(t), 0.08))
            r = np.full(4, 0.4, f32)
        for val, wgt in terms:
            r = (r + (val.astype(f32) * f32(wgt)).astype(f32)).astype(f32)
        return r
    if op in ("phi_loop", "phi_swap"):
        n = 3 + (int(np.trunc(f32(a[0] * f32(8.0)))) & 3)
        acc, oth = a.copy(), b.copy()
        for _ in range(n):
            val = ((acc * f32(0.5)).astype(f32) + (oth * c).astype(f32)).astype(f32)
            if op == "phi_swap":
                acc, oth = oth, (val + f32(0)).astype(f32)
            else:
                acc, oth = (val + f32(0)).astype(f32), (oth + f32(0)).astype(f32)
        return ((acc + oth).astype(f32) * f32(0.25)).astype(f32)
    raise ValueError(op)


@pytest.mark.parametrize("op", shaders.EXT_UNIT_OPS)
def test_extended_op_rejected_by_default_and_matches_numpy_when_enabled(vor, op):
    """opcodes beyond Appendix B are refused exactly like the reference refuses them
    (CompileFunction == NULL, spirv_compile.cpp:1734,1888) unless the extended option is on"""
    setopt = vor.lib.vor_set_option
    setopt.argtypes = [C.c_char_p, C.c_int64]
    with pytest.raises(abi.BackendError):
        vor.CompileFunction(shaders.vs_unit(op))
    assert setopt(b"extended_spirv", 1) == 0
    try:
        verts, ubo = unit_inputs(3)
        global _UBO_FOR_EXT
        _UBO_FOR_EXT = ubo
        verts[::3, 4:8] = verts[::3, 0:4]    # equal operands for the (in)equality tests
        mod, st, keep = unit_state(vor, op, verts, ubo)
        entry = vor.GetFuncPointer(mod, "main")
        run = vor.lib.vor_run_vertex
        run.argtypes = [C.POINTER(abi.DrawState), C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
        out = (C.c_float * 44)()
        for i in range(verts.shape[0]):
            assert run(C.byref(st), entry, i, out) == 0
            got = np.frombuffer(out, dtype=f32)[4:8].copy()    # bit patterns (isub produces NaN payloads)
            a, b, c = verts[i, 0:4], verts[i, 4:8], verts[i, 8:12]
            exp = np.asarray(_expected_ext(op, a, b, c), dtype=f32)
            if op in shaders.APPROX_EXT_OPS:    # libm vs numpy: close, not bit-reproducible by construction
                assert np.allclose(got, exp, rtol=1e-5, atol=2e-6), (op, i, got, exp)
                continue
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (op, i, got, exp)
        vor.DestroyFunction(mod)
    finally:
        setopt(b"extended_spirv", 0)


def test_explicit_lod_sample_equals_implicit(vor):
    """extended mode: OpImageSampleExplicitLod ignores its Lod operand (mip 0, like every sample of the
    reference's sampler), refused by default like any opcode outside Appendix B"""
    from harness import scenes
    setopt = vor.lib.vor_set_option
    setopt.argtypes = [C.c_char_p, C.c_int64]
    with pytest.raises(abi.BackendError):
        vor.CompileFunction(shaders.fs_texture(True))
    assert setopt(b"extended_spirv", 1) == 0
    try:
        sc = scenes.c2_cube(160, 100, tex_size=32)
        sc.draws[0].pipe.fs = shaders.fs_texture(True)
        got, gd = scenes.render(vor, sc)
        want, wd = scenes.render(vor, scenes.c2_cube(160, 100, tex_size=32))
        assert np.array_equal(got, want) and np.array_equal(gd, wd)
        assert (got != 0xCD).any()
    finally:
        setopt(b"extended_spirv", 0)


def test_fragment_interpolation_kat(vor):
    """FS input = ((b0*v0 + b1*v1) + b2*v2) + bw*0 per component (spirv_compile.cpp:629-643,2196-2214)."""
    mod = vor.CompileFunction(shaders.fs_color())
    entry = vor.GetFuncPointer(mod, "main")
    run = vor.lib.vor_run_fragment
    run.argtypes = [C.POINTER(abi.DrawState), C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                    C.POINTER(C.c_float)]
    rng = np.random.default_rng(5)
    st = abi.DrawState()
    for _ in range(50):
        tri = rng.uniform(-4, 4, size=(3, 44)).astype(f32)
        bary = rng.uniform(0, 1, size=4).astype(f32)
        bary[3] = 0
        out = (C.c_float * 4)()
        assert run(C.byref(st), entry, bary.ctypes.data_as(C.POINTER(C.c_float)),
                   tri.ctypes.data_as(C.POINTER(C.c_float)), out) == 0
        exp = []
        for comp in range(4):
            v = [tri[k, 4 + comp] for k in range(3)]
            acc = f32(f32(bary[0] * v[0]) + f32(bary[1] * v[1]))
            acc = f32(acc + f32(bary[2] * v[2]))
            acc = f32(acc + f32(bary[3] * f32(0.0)))
            exp.append(acc)
        assert np.array_equal(np.array(out[:], dtype=f32).view(np.uint32), np.array(exp, dtype=f32).view(np.uint32))


def test_vertex_output_padding_rules(vor):
    """vec<4 outputs repeat .x, scalars splat, ints are bit-cast (spirv_compile.cpp:2057-2092)."""
    mod = vor.CompileFunction(shaders.vs_kitchen_sink())
    entry = vor.GetFuncPointer(mod, "main")
    rng = np.random.default_rng(9)
    verts = np.zeros(6, dtype=[("pos", f32, 4), ("nrm", f32, 3), ("w", f32), ("flag", np.int32)])
    verts["pos"] = rng.uniform(-1, 1, (6, 4))
    verts["nrm"] = rng.uniform(-1, 1, (6, 3))
    verts["w"] = 0.75
    verts["flag"] = 7
    vb = verts.view(np.uint8)
    ubo = rng.uniform(-1, 1, 36).astype(f32)
    pl = abi.Pipeline()
    for loc, (fmt, off) in enumerate([(abi.FMT_R32G32B32A32_SFLOAT, 0), (abi.FMT_R32G32B32_SFLOAT, 16),
                                      (abi.FMT_R32_SFLOAT, 28), (abi.FMT_R32_SINT, 32)]):
        pl.vattrs[loc].format, pl.vattrs[loc].stride, pl.vattrs[loc].offset = fmt, 36, off
    st = abi.DrawState()
    st.vbs[0].buffer = abi.make_buffer(vb)
    st.pipeline = C.pointer(pl)
    arr = (abi.Binding * 1)()
    arr[0].set, arr[0].binding = 1, 2
    arr[0].buffer = abi.make_buffer(ubo)
    st.bindings = C.cast(arr, C.POINTER(abi.Binding))
    st.num_bindings = 1
    push = np.array([0.5, -0.25, 2.0, 1.5], dtype=f32)
    C.memmove(st.pushconsts, push.ctypes.data, 16)
    run = vor.lib.vor_run_vertex
    run.argtypes = [C.POINTER(abi.DrawState), C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
    out = (C.c_float * 44)()
    assert run(C.byref(st), entry, 5, out) == 0
    o = np.array(out[:], dtype=f32).reshape(11, 4)
    assert np.array_equal(o[0], verts["pos"][5])                 # gl_Position = pos (vertex 5)
    assert o[2][3] == o[2][0]                                    # vec3 @1 padded with .x
    assert o[3][0] == o[3][1] == o[3][2] == o[3][3]              # float @2 splat
    flag_bits = o[4].view(np.int32)
    assert (flag_bits == flag_bits[0]).all()
    # ((7*3) & 0xff) >> 1 [OpShiftLeftLogical is a right shift in the reference] + gl_VertexIndex(5)
    assert flag_bits[0] == ((7 * 3) & 0xFF) // 2 + 5
    assert o[5][2] == o[5][0] and o[5][3] == o[5][0]             # vec2 @4 padded with .x
    assert o[5][0] == push[1] and o[5][1] == ubo[34]             # shuffle(k, s, 1, 6) = (k.y, s.z)
    assert np.array_equal(o[9], push)                            # mat4 @5..8: last column = k


@pytest.mark.parametrize("in_callee", [False, True])
def test_discard_leaves_colour_and_depth_untouched(vor, in_callee):
    """extended mode, OpKill (no reference counterpart: CompileFunction asserts on it, spirv_compile.cpp:1888):
    refused by default; when enabled a discarded fragment writes neither colour nor depth, every other pixel is
    what the same scene gives without the discard. Known answer: the discard tests the interpolated red against
    0.5, and the scene without it tells the red every fragment has."""
    from harness import scenes
    setopt = vor.lib.vor_set_option
    setopt.argtypes = [C.c_char_p, C.c_int64]
    with pytest.raises(abi.BackendError):
        vor.CompileFunction(shaders.fs_color_kill(0.5, in_callee))
    assert setopt(b"extended_spirv", 1) == 0
    try:
        def scene(kill):
            sc = scenes.random_triangles(160, 96, 1, 5, max_size=1.9, perspective=False, offscreen=0.0)
            if kill:
                sc.draws[0].pipe.fs = shaders.fs_color_kill(0.5, in_callee)
            return sc
        plain_c, plain_d = scenes.render(vor, scene(False))
        kill_c, kill_d = scenes.render(vor, scene(True))
        covered = plain_d != np.float32(1.0)              # the one triangle's pixels (cleared depth is 1.0)
        red = plain_c[..., 2]                             # BGRA bytes: byte(red * 255)
        assert covered.sum() > 1000
        killed = covered & (red < 127)
        kept = covered & (red > 127)
        assert killed.sum() > 100 and kept.sum() > 100
        clear_c = scenes.render(vor, scenes.random_triangles(160, 96, 1, 5, max_size=0.0))[0][0, 0]
        assert (kill_c[killed] == clear_c).all() and (kill_d[killed] == np.float32(1.0)).all()
        assert np.array_equal(kill_c[kept], plain_c[kept]) and np.array_equal(kill_d[kept], plain_d[kept])
        assert np.array_equal(kill_c[~covered], plain_c[~covered])
    finally:
        setopt(b"extended_spirv", 0)
