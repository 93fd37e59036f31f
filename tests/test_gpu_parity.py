"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): coverage, depth-test outcomes and integer/UNORM framebuffer writes
bit-exact; shaded colour within 1 UNORM8 LSB — in practice every scene here is required to be
byte-identical because the shader arithmetic is IEEE-exact in both (no sin/cos/pow in these scenes).
"""
import ctypes as C

import numpy as np
import pytest

from harness import abi, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "ordered"], autouse=True)
def raster_path(request, gpu):
    """Every scene runs through both tile back ends: "auto" picks the visibility-resolve kernel when the
    pass is order independent, "ordered" forces the in-order kernel (exact for any state)."""
    import ctypes as C
    gpu.lib.vb200_set_option.argtypes = [C.c_char_p, C.c_int64]
    assert gpu.lib.vb200_set_option(b"raster_path", 1 if request.param == "ordered" else 0) == 0
    yield request.param
    gpu.lib.vb200_set_option(b"raster_path", 0)


def _check(gpu, vor, sc, exact=True):
    c_gpu, d_gpu = scenes.render(gpu, sc)
    c_cpu, d_cpu = scenes.render(vor, sc)
    def where(mask):
        ys, xs = np.nonzero(mask)
        return f"x {xs.min()}..{xs.max()}, y {ys.min()}..{ys.max()}, first at ({xs[0]},{ys[0]})"

    if d_cpu is not None:
        badmask = d_gpu.view(np.uint32) != d_cpu.view(np.uint32)
        bad = badmask.sum()
        assert bad == 0, (f"{sc.name}: {bad} depth words differ ({where(badmask.reshape(sc.height, sc.width))}); "
                          f"gpu {d_gpu[badmask][:4]} cpu {d_cpu[badmask][:4]}")
    diff = np.abs(c_gpu.astype(np.int16) - c_cpu.astype(np.int16))
    if exact:
        badmask = (diff > 0).any(-1)
        assert diff.max() == 0, (f"{sc.name}: {badmask.sum()} pixels differ (max {diff.max()} LSB; "
                                 f"{where(badmask.reshape(sc.height, sc.width))}); gpu {c_gpu[badmask][:3].tolist()} "
                                 f"cpu {c_cpu[badmask][:3].tolist()}")
    else:
        assert diff.max() <= 1, f"{sc.name}: max colour difference {diff.max()} LSB"
    return float((diff > 0).any(-1).mean())


@pytest.mark.parametrize("name,builder", [
    ("c1", lambda: scenes.c1_triangle()),
    ("c1_odd", lambda: scenes.c1_triangle(333, 211)),
    ("c2", lambda: scenes.c2_cube(960, 540)),
    ("c3", lambda: scenes.c3_mesh(960, 540, 250, 125)),
    ("c4", lambda: scenes.c4_particles(640, 360, 20000)),
    ("c5", lambda: scenes.c5_textured(960, 540, 250, 125, tex_size=256)),
])
def test_config_scenes(gpu, vor, name, builder):
    _check(gpu, vor, builder())


@pytest.mark.parametrize("seed", range(6))
def test_random_depth(gpu, vor, seed):
    ops = [abi.CMP_LESS, abi.CMP_LEQUAL, abi.CMP_GREATER, abi.CMP_GEQUAL, abi.CMP_EQUAL, abi.CMP_NOTEQUAL]
    sc = scenes.random_triangles(400, 300, 300, seed, depth_op=ops[seed], depth_write=(seed % 3 != 2))
    if ops[seed] in (abi.CMP_GREATER, abi.CMP_GEQUAL):
        sc.clear_depth = 0.0
    if ops[seed] == abi.CMP_EQUAL:
        sc.clear_depth = 0.5
    _check(gpu, vor, sc)


@pytest.mark.parametrize("seed", range(3))
def test_random_blend(gpu, vor, seed):
    factors = [(abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA), (abi.BF_ONE, abi.BF_ONE), (abi.BF_ZERO, abi.BF_SRC_ALPHA)]
    sc = scenes.random_triangles(400, 300, 400, 100 + seed, blend=(*factors[seed], abi.BLEND_ADD),
                                 depth_op=abi.CMP_ALWAYS, has_depth=False)
    _check(gpu, vor, sc)


def test_cull_strip_index16(gpu, vor):
    _check(gpu, vor, scenes.random_triangles(320, 200, 60, 8, topology=abi.TOPO_STRIP, depth_op=abi.CMP_LEQUAL))
    _check(gpu, vor, scenes.random_triangles(320, 200, 200, 9, index_type=abi.INDEX_U16, cull=abi.CULL_BACK,
                                             front=abi.FRONT_CW))
    _check(gpu, vor, scenes.random_triangles(320, 200, 200, 10, index_type=abi.INDEX_U32, cull=abi.CULL_FRONT))


def test_depth_test_without_write_and_never(gpu, vor):
    for op in (abi.CMP_LESS, abi.CMP_GEQUAL, abi.CMP_NEVER, abi.CMP_NOTEQUAL, abi.CMP_EQUAL):
        sc = scenes.random_triangles(300, 200, 200, 40 + op, depth_op=op, depth_write=False)
        sc.clear_depth = 0.5
        _check(gpu, vor, sc)
    sc = scenes.random_triangles(300, 200, 200, 50, depth_op=abi.CMP_ALWAYS, depth_write=True)
    _check(gpu, vor, sc)


def test_large_and_small_mix(gpu, vor):
    """full-screen triangles (row-swept by the whole CTA) interleaved with tiny ones"""
    big = scenes.random_triangles(500, 300, 12, 60, max_size=1.5)
    small = scenes.random_triangles(500, 300, 3000, 61, max_size=0.02)
    big.draws += small.draws
    _check(gpu, vor, big)


def test_two_draws_accumulate_depth(gpu, vor):
    a = scenes.random_triangles(300, 200, 150, 70)
    b = scenes.random_triangles(300, 200, 150, 71, depth_op=abi.CMP_LEQUAL)
    a.draws += b.draws
    _check(gpu, vor, a)


def test_no_clear_loads_host_contents(gpu, vor):
    """loadOp LOAD: attachments keep what the host put there (coherent memory semantics)."""
    sc = scenes.random_triangles(200, 120, 50, 3)
    sc.clear_color = None
    sc.clear_depth = None
    _check(gpu, vor, sc)


def _unit_scene(op, seed=0):
    """Random triangles whose colour is the result of one SPIR-V op evaluated in the vertex stage."""
    from harness import shaders
    rng = np.random.default_rng(seed)
    n = 40
    v = np.zeros((n * 3, 12), dtype=np.float32)
    c = rng.uniform(-0.9, 0.9, size=(n, 1, 2))
    v[:, 0:2] = (c + rng.uniform(-0.3, 0.3, size=(n, 3, 2))).reshape(-1, 2)
    v[:, 2] = 0.5
    v[:, 3] = 1.0
    v[:, 4:12] = rng.uniform(0.05, 1.0, size=(n * 3, 8))
    if op in ("fsub", "fneg", "reflect3"):
        v[:, 4:8] *= 0.3
    ubo = rng.uniform(0.0, 0.4, size=32).astype(np.float32)
    pipe = scenes.PipelineDesc(shaders.vs_unit(op), shaders.fs_color(),
                               [(0, abi.FMT_R32G32B32A32_SFLOAT, 48, 0, 0), (1, abi.FMT_R32G32B32A32_SFLOAT, 48, 16, 0),
                                (2, abi.FMT_R32G32B32A32_SFLOAT, 48, 32, 0)])
    d = scenes.Draw(pipe, n * 3, vbs=[(v, 0)], ubos=[(0, 0, ubo, 0)])
    return scenes.Scene(f"unit_{op}", 256, 160, [d], depth=False)


@pytest.mark.parametrize("op", [o for o in (lambda sh: sh.UNIT_OPS + sh.MEM_UNIT_OPS)(__import__("harness.shaders", fromlist=["UNIT_OPS"]))])
def test_spirv_ops_match_oracle(gpu, vor, op):
    # sin/cos/pow are libm in the oracle and approx units on the GPU: covered by the 1-LSB colour bar
    _check(gpu, vor, _unit_scene(op), exact=op not in ("sin", "cos", "pow"))


@pytest.mark.parametrize("op", [o for o in __import__("harness.shaders", fromlist=["EXT_UNIT_OPS"]).EXT_UNIT_OPS])
def test_extended_spirv_ops(gpu, vor, op):
    """opcodes beyond the reference's subset: refused by default (as CompileFunction refuses them), and with
    the "extended_spirv" option the PTX they lower to agrees with the CPU interpreter bit for bit"""
    from harness import shaders
    gset, oset = gpu.lib.vb200_set_option, vor.lib.vor_set_option
    gset.argtypes = oset.argtypes = [C.c_char_p, C.c_int64]
    with pytest.raises(abi.BackendError):
        gpu.CompileFunction(shaders.vs_unit(op))
    assert gset(b"extended_spirv", 1) == 0 and oset(b"extended_spirv", 1) == 0
    try:
        sc = _unit_scene(op)
        v = sc.draws[0].vbs[0][0]
        v[::3, 8:12] = v[::3, 4:8]          # equal operands for the (in)equality tests
        v[1::3, 4:8] = -v[1::3, 4:8] * 7    # negatives / magnitudes > 1 for floor, fract, fabs, ftos
        # (the transcendental ones: libm in the oracle, the special-function unit here — the 1-LSB colour bar)
        _check(gpu, vor, sc, exact=op not in shaders.APPROX_EXT_OPS)
    finally:
        gset(b"extended_spirv", 0)
        oset(b"extended_spirv", 0)


@pytest.mark.parametrize("variant", ["depth", "blend", "callee"])
def test_discard_in_the_fragment_shader(gpu, vor, variant):
    """extended mode, OpKill: a discarded fragment writes neither colour nor depth. The colour a pixel ends up with
    is that of the last fragment that passed and was KEPT, so such shaders always run on the in-order tile kernel
    (also when the pass would otherwise be order independent)."""
    from harness import shaders
    gset, oset = gpu.lib.vb200_set_option, vor.lib.vor_set_option
    gset.argtypes = oset.argtypes = [C.c_char_p, C.c_int64]
    with pytest.raises(abi.BackendError):
        gpu.CompileFunction(shaders.fs_color_kill())
    assert gset(b"extended_spirv", 1) == 0 and oset(b"extended_spirv", 1) == 0
    try:
        blend = (abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, 0) if variant == "blend" else None
        sc = scenes.random_triangles(300, 200, 300, 17, blend=blend)
        sc.draws[0].pipe.fs = shaders.fs_color_kill(0.45, in_callee=(variant == "callee"))
        _check(gpu, vor, sc)
        gpu.lib.vb200_last_tile_kernel.restype = C.c_char_p
        assert gpu.lib.vb200_last_tile_kernel() == b"vb200_k_tile_ordered"
    finally:
        gset(b"extended_spirv", 0)
        oset(b"extended_spirv", 0)


def test_sample_with_explicit_lod(gpu, vor):
    """extended mode, OpImageSampleExplicitLod (textureLod): the level is ignored — the sampler reads mip 0 for
    every instruction, as the reference's does — so the frame equals the one OpImageSampleImplicitLod gives"""
    from harness import shaders
    gset, oset = gpu.lib.vb200_set_option, vor.lib.vor_set_option
    gset.argtypes = oset.argtypes = [C.c_char_p, C.c_int64]
    with pytest.raises(abi.BackendError):
        gpu.CompileFunction(shaders.fs_texture(True))
    assert gset(b"extended_spirv", 1) == 0 and oset(b"extended_spirv", 1) == 0
    try:
        sc = scenes.c2_cube(480, 270, tex_size=64)
        sc.draws[0].pipe.fs = shaders.fs_texture(True)
        _check(gpu, vor, sc)
        want, _ = scenes.render(gpu, scenes.c2_cube(480, 270, tex_size=64))
        got, _ = scenes.render(gpu, sc)
        assert np.array_equal(got, want)
    finally:
        gset(b"extended_spirv", 0)
        oset(b"extended_spirv", 0)


def test_resolve_without_slot_keys(gpu, vor):
    """draws with >= 2^24 triangles cannot carry the record slot in the visibility key; the option forces
    that code path (phase B gathers the winner from global memory) on ordinary scenes"""
    setopt = gpu.lib.vb200_set_option
    setopt.argtypes = [C.c_char_p, C.c_int64]
    assert setopt(b"slot_keys", 0) == 0
    try:
        _check(gpu, vor, scenes.c3_mesh(480, 270, 120, 60))
        for seed, op in enumerate([abi.CMP_LESS, abi.CMP_LEQUAL, abi.CMP_GREATER, abi.CMP_GEQUAL]):
            sc = scenes.random_triangles(300, 200, 200, 60 + seed, depth_op=op)
            if op in (abi.CMP_GREATER, abi.CMP_GEQUAL):
                sc.clear_depth = 0.0
            _check(gpu, vor, sc)
        _check(gpu, vor, scenes.random_triangles(300, 200, 200, 70, depth_op=abi.CMP_ALWAYS, depth_write=False))
    finally:
        setopt(b"slot_keys", 1)


def test_many_draws_are_batched_and_keep_submission_order(gpu, vor):
    """consecutive draws with one pipeline and one set of bindings are rasterised as ONE batch (one vertex,
    setup and tile pass); triangle ids keep the submission order across draws, so blended and equal-depth
    results equal the reference's draw-by-draw replay (cmd_exec.cpp:129-142)"""
    import ctypes as C
    # indexed mesh cut into 40 draws (together they reference every vertex: one shared vertex span)
    sc = scenes.split_draws(scenes.c3_mesh(480, 270, 120, 60), 40)
    gpu.reset_stats()
    _check(gpu, vor, sc)
    st = gpu.stats()
    assert st["draws"] == 40
    assert st["kernel_launches"] <= 4, st    # vertex + setup (+ sort) + tiles for all 40 draws
    # a few small indexed draws out of a large buffer: host-measured spans, draws submitted out of order
    _check(gpu, vor, scenes.split_draws(scenes.c3_mesh(480, 270, 120, 60), 40, order=[31, 3, 17, 4, 5]))
    # non-indexed draws, blended (order dependent) and depth-tested with ties, shuffled submission order
    rng = np.random.default_rng(5)
    for kw in (dict(blend=(abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, 0), depth_op=abi.CMP_ALWAYS, depth_write=False),
               dict(depth_op=abi.CMP_LEQUAL), dict(depth_op=abi.CMP_LESS), dict(depth_op=abi.CMP_NOTEQUAL)):
        base = scenes.random_triangles(300, 200, 240, 33, **kw)
        _check(gpu, vor, scenes.split_draws(base, 24, order=list(rng.permutation(24))))
    # indexed u16 draws through a shuffled vertex buffer
    base = scenes.random_triangles(300, 200, 120, 34, index_type=abi.INDEX_U16)
    _check(gpu, vor, scenes.split_draws(base, 12, order=list(rng.permutation(12))))
    # every draw by itself gives the same image
    setopt = gpu.lib.vb200_set_option
    setopt.argtypes = [C.c_char_p, C.c_int64]
    assert setopt(b"batch_draws", 0) == 0
    try:
        _check(gpu, vor, scenes.split_draws(scenes.c3_mesh(480, 270, 120, 60), 7))
    finally:
        setopt(b"batch_draws", 1)


def test_state_changes_split_batches(gpu, vor):
    """draws whose pipeline, bindings or targets differ cannot share a batch; the sequence must still give the
    reference's image (second pipeline: other cull mode and depth op on the same targets)"""
    import dataclasses
    a = scenes.random_triangles(320, 240, 150, 41, depth_op=abi.CMP_LESS)
    b2 = scenes.random_triangles(320, 240, 150, 42, depth_op=abi.CMP_GEQUAL, cull=abi.CULL_BACK)
    c = scenes.random_triangles(320, 240, 150, 43, depth_op=abi.CMP_LESS,
                                blend=(abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, 0))
    parts = [scenes.split_draws(x, 5).draws for x in (a, b2, c)]
    draws = [d for trio in zip(*parts) for d in trio] + parts[0][:2]
    _check(gpu, vor, dataclasses.replace(a, name="interleaved_pipelines", draws=draws))


def test_kitchen_sink_shaders(gpu, vor):
    """function calls, loops, branches, push constants, UBO at (set 1, binding 2), int/flat and matrix
    varyings, through both stages"""
    from harness import shaders
    rng = np.random.default_rng(11)
    n = 60
    verts = np.zeros(n * 3, dtype=[("pos", np.float32, 4), ("nrm", np.float32, 3), ("w", np.float32), ("flag", np.int32)])
    c = rng.uniform(-0.9, 0.9, size=(n, 1, 2))
    verts["pos"][:, 0:2] = (c + rng.uniform(-0.4, 0.4, size=(n, 3, 2))).reshape(-1, 2)
    verts["pos"][:, 2] = rng.uniform(0.1, 0.9, n * 3)
    verts["pos"][:, 3] = 1.0
    verts["nrm"] = rng.uniform(-1, 1, size=(n * 3, 3))
    verts["w"] = rng.uniform(0, 1, n * 3)
    verts["flag"] = np.repeat(rng.integers(0, 12, n), 3)
    vb = np.ascontiguousarray(verts).view(np.uint8)
    ubo = rng.uniform(-0.5, 0.5, size=36).astype(np.float32)
    push = np.array([0.3, 0.6, 0.1, 0.9], dtype=np.float32).tobytes()
    pipe = scenes.PipelineDesc(shaders.vs_kitchen_sink(), shaders.fs_kitchen_sink(),
                               [(0, abi.FMT_R32G32B32A32_SFLOAT, 36, 0, 0), (1, abi.FMT_R32G32B32_SFLOAT, 36, 16, 0),
                                (2, abi.FMT_R32_SFLOAT, 36, 28, 0), (3, abi.FMT_R32_SINT, 36, 32, 0)],
                               depth_op=abi.CMP_LESS, depth_write=True)
    d = scenes.Draw(pipe, n * 3, vbs=[(vb, 0)], ubos=[(1, 2, ubo, 0)], push=push)
    _check(gpu, vor, scenes.Scene("kitchen_sink", 320, 200, [d], depth=True))


def test_sampler_matches_oracle(gpu, vor):
    rng = np.random.default_rng(3)
    tex = rng.integers(0, 256, size=(64, 32, 4), dtype=np.uint8)
    im = abi.make_image(tex, 32, 64, abi.FMT_R8G8B8A8_UNORM)
    uv = rng.uniform(0, 6, size=(20000, 2)).astype(np.float32)
    uv[:8] = [[0, 0], [1, 1], [0.999999, 0.5], [5.0, 0.25]] * 2
    a, b = gpu.sample(im, uv), vor.sample(im, uv)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    cube = rng.integers(0, 256, size=(6, 16, 16, 4), dtype=np.uint8)
    cim = abi.make_image(cube, 16, 16, abi.FMT_R8G8B8A8_UNORM, layers=6)
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    a, b = gpu.sample(cim, d, cube=True), vor.sample(cim, d, cube=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_unorm8_conversion_is_the_ieee_quotient(gpu):
    """texel conversion float(byte) / 255.0f (texture_sampling.cpp:121-133) is computed without a division on
    the device (one Newton step); sampling exactly at the texel origins returns the conversion itself
    (weights 1 and 0), so all 256 byte values are compared with the IEEE single-precision quotient"""
    tex = np.arange(256, dtype=np.uint8).repeat(4).reshape(1, 256, 4).repeat(4, axis=0).copy()    # 256 x 4, RGBA = (b,b,b,b)
    tex[:, :, 1] = 255 - tex[:, :, 1]
    tex[:, :, 2] = (tex[:, :, 2].astype(np.uint16) * 7 % 256).astype(np.uint8)
    img = abi.make_image(tex, 256, 4, abi.FMT_R8G8B8A8_UNORM)
    uvw = np.zeros((256, 2), dtype=np.float32)
    uvw[:, 0] = np.arange(256, dtype=np.float32) / np.float32(256.0)    # exact: u * 256 is the integer texel column
    out = gpu.sample(img, uvw)
    want = tex[0].astype(np.float32) / np.float32(255.0)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_sampler_block_compressed_and_r8(gpu, vor):
    """BC2/BC3 decode and the 1-byte-per-pixel linear path on the device texture unit"""
    rng = np.random.default_rng(11)
    uv = rng.uniform(0.0, 4.0, size=(30000, 2)).astype(np.float32)
    for fmt in (abi.FMT_BC2_UNORM_BLOCK, abi.FMT_BC3_UNORM_BLOCK):
        data = scenes.bc_blocks(rng, 64, 32)
        im = abi.make_image(data, 64, 32, fmt, bpp=1)
        a, b = gpu.sample(im, uv), vor.sample(im, uv)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), fmt
        # an image bound at an odd byte offset (alignment 1, images.cpp:51-56): unaligned block loads
        shifted = np.zeros(data.size + 3, np.uint8)
        shifted[3:] = data
        im2 = abi.make_image(shifted[3:], 64, 32, fmt, bpp=1)
        c = gpu.sample(im2, uv)
        assert np.array_equal(c.view(np.uint32), b.view(np.uint32)), fmt
    # 1-byte texels are fetched 4 bytes at a time (texture_sampling.cpp:121-133): the last texels read past the
    # image. The bytes behind it are mirrored when the allocation is known to extend that far (registered).
    r8 = rng.integers(0, 256, size=(32 * 16 + 4,), dtype=np.uint8)
    L = _copy_api(gpu)
    gpu.check(L.vb200_mem_register(r8.ctypes.data, r8.nbytes), "mem_register")
    try:
        im = abi.make_image(r8, 32, 16, abi.FMT_R8_UNORM, bpp=1)
        uv[:, 1] = uv[:, 1] % 0.9
        a, b = gpu.sample(im, uv), vor.sample(im, uv)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    finally:
        L.vb200_mem_unregister(r8.ctypes.data)


@pytest.mark.parametrize("fmt", [abi.FMT_BC2_UNORM_BLOCK, abi.FMT_BC3_UNORM_BLOCK])
def test_block_compressed_textured_scene(gpu, vor, fmt):
    """the C2 cube with its texture stored as BC2 / BC3 blocks"""
    sc = scenes.c2_cube(640, 360)
    d = sc.draws[0]
    s, b, _, _, _, _, _, layers = d.textures[0]
    d.textures = [(s, b, scenes.bc_blocks(np.random.default_rng(21), 128, 64), 128, 64, fmt, 1, layers)]
    _check(gpu, vor, sc)


def _copy_api(gpu):
    L = gpu.lib
    L.vb200_copy_buffer.argtypes = [C.POINTER(abi.Buffer), C.c_uint64, C.POINTER(abi.Buffer), C.c_uint64, C.c_uint64]
    L.vb200_copy_buffer_to_image.argtypes = [C.POINTER(abi.Buffer), C.c_uint64, C.POINTER(abi.Image), C.c_uint32,
                                             C.c_uint32]
    L.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mem_unregister.argtypes = [C.c_void_p]
    L.vb200_mem_set_device_local.argtypes = [C.c_void_p, C.c_int]
    return L


def test_copy_buffer_and_copy_buffer_to_image(gpu, vor):
    """vkCmdCopyBuffer / vkCmdCopyBufferToImage replayed on the device (cmd_exec.cpp:143-182): host memory
    ends up exactly as the reference's memcpy leaves it, and the copy is ordered with the draws"""
    L = _copy_api(gpu)
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, size=5000, dtype=np.uint8)
    dst = np.full(6000, 7, np.uint8)
    sb, db = abi.make_buffer(src), abi.make_buffer(dst)
    gpu.check(L.vb200_copy_buffer(C.byref(sb), 100, C.byref(db), 1000, 3000), "copy_buffer")
    gpu.flush()
    want = np.full(6000, 7, np.uint8)
    want[1000:4000] = src[100:3100]
    assert np.array_equal(dst, want)
    assert L.vb200_copy_buffer(C.byref(sb), 4000, C.byref(db), 0, 2000) != 0    # source range outside the buffer

    # a 2-layer image with 3 mips (32x16, 16x8, 8x4): CalcSubresourceByteOffset (precompiled.cpp:3-36)
    chain = 32 * 16 * 4 + 16 * 8 * 4 + 8 * 4 * 4
    img = np.zeros(2 * chain, np.uint8)
    im = abi.make_image(img, 32, 16, abi.FMT_R8G8B8A8_UNORM, layers=2, mips=3)
    gpu.check(L.vb200_copy_buffer_to_image(C.byref(sb), 16, C.byref(im), 1, 1), "copy_buffer_to_image")
    gpu.flush()
    want = np.zeros(2 * chain, np.uint8)
    offs = chain + 32 * 16 * 4
    want[offs:offs + 16 * 8 * 4] = src[16:16 + 16 * 8 * 4]
    assert np.array_equal(img, want)


def test_device_local_texture_filled_by_copy(gpu, vor):
    """a texture in DEVICE_LOCAL memory: filled by a device copy from a staging buffer, sampled by the
    draw, never read from or written to its host shadow"""
    L = _copy_api(gpu)
    sc = scenes.c2_cube(640, 360)
    want_c, want_d = scenes.render(vor, sc)
    d = sc.draws[0]
    s, b, tex, tw, th, fmt, bpp, layers = d.textures[0]
    staging = np.ascontiguousarray(tex).reshape(-1).copy()
    shadow = np.zeros(staging.size + 16, np.uint8)    # "device local": the host side stays zero
    gpu.check(L.vb200_mem_register(shadow.ctypes.data, shadow.nbytes), "mem_register")
    try:
        gpu.check(L.vb200_mem_set_device_local(shadow.ctypes.data, 1), "set_device_local")
        sb = abi.make_buffer(staging)
        im = abi.make_image(shadow, tw, th, fmt, bpp=bpp, layers=layers)
        gpu.check(L.vb200_copy_buffer_to_image(C.byref(sb), 0, C.byref(im), 0, 0), "copy_buffer_to_image")
        gpu.flush()                                   # a later submit still finds the texture in HBM
        d.textures = [(s, b, shadow[:staging.size].reshape(tex.shape), tw, th, fmt, bpp, layers)]
        for _ in range(2):
            got_c, got_d = scenes.render(gpu, sc)
            assert np.array_equal(got_c, want_c) and np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
        assert not shadow.any()
    finally:
        L.vb200_mem_unregister(shadow.ctypes.data)


def test_cube_map_scene(gpu, vor):
    from harness import shaders
    rng = np.random.default_rng(4)
    cube = rng.integers(0, 256, size=(6, 32, 32, 4), dtype=np.uint8)
    n = 30
    v = np.zeros((n * 3, 8), dtype=np.float32)
    c = rng.uniform(-0.8, 0.8, size=(n, 1, 2))
    v[:, 0:2] = (c + rng.uniform(-0.5, 0.5, size=(n, 3, 2))).reshape(-1, 2)
    v[:, 2], v[:, 3] = 0.5, 1.0
    v[:, 4:7] = rng.normal(size=(n * 3, 3))
    pipe = scenes.PipelineDesc(shaders.vs_pos_dir(), shaders.fs_cube(),
                               [(0, abi.FMT_R32G32B32A32_SFLOAT, 32, 0, 0), (1, abi.FMT_R32G32B32_SFLOAT, 32, 16, 0)])
    d = scenes.Draw(pipe, n * 3, vbs=[(v, 0)], textures=[(0, 1, cube, 32, 32, abi.FMT_R8G8B8A8_UNORM, 4, 6)])
    _check(gpu, vor, scenes.Scene("cube", 300, 200, [d], depth=False))


def test_full_size_c3_properties(gpu):
    """At BASELINE.json's full size the oracle image is pinned by tests/golden (test_golden.py); here the
    size-independent properties: idempotence (same frame twice -> identical bytes), both tile back ends
    agree, and every pixel the draw did not touch still holds the clear value."""
    import ctypes as C
    sc = scenes.c3_mesh()
    c1, d1 = scenes.render(gpu, sc)
    c2, d2 = scenes.render(gpu, sc)
    assert np.array_equal(c1, c2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    untouched = d1 == np.float32(1.0)
    assert (c1[untouched] == np.array([51, 51, 51, 255], dtype=np.uint8)).all()
    assert (d1 <= np.float32(1.0)).all() and untouched.mean() < 0.5


def test_many_pairs_per_triangle(gpu, vor):
    """40 near-full-screen triangles at 2048x2048: every triangle is appended to thousands of tile lists
    (the CTA-wide walk of the binning pass)"""
    sc = scenes.random_triangles(2048, 2048, 40, 80, max_size=1.6, offscreen=0.0)
    _check(gpu, vor, sc)


@pytest.mark.parametrize("cap", [1, 3, 40])
def test_tile_list_overflow_falls_back_to_range_scan(gpu, vor, cap):
    """A tile that receives more triangles than its list holds is rasterised from the packed per-triangle
    tile ranges instead (exact and in submission order, no host round trip). `tile_list_cap` shrinks the
    lists so that most / some tiles take that path: small and large triangles, one and several rounds of
    256 triangles per tile, both tile back ends."""
    gpu.lib.vb200_set_option.argtypes = [C.c_char_p, C.c_int64]
    assert gpu.lib.vb200_set_option(b"tile_list_cap", cap) == 0
    try:
        _check(gpu, vor, scenes.c3_mesh(640, 360, 250, 125))              # ~10 triangles per tile... up to hundreds
        _check(gpu, vor, scenes.c3_mesh(96, 64, 250, 125))                # thousands of triangles per tile: many rounds
        _check(gpu, vor, scenes.c4_particles(320, 200, 6000))
        big = scenes.random_triangles(500, 300, 12, 60, max_size=1.5)
        big.draws += scenes.random_triangles(500, 300, 3000, 61, max_size=0.02, depth_op=abi.CMP_LEQUAL).draws
        _check(gpu, vor, big)
    finally:
        gpu.lib.vb200_set_option(b"tile_list_cap", 0)


def _render_with(be, sc, color0=None, depth0=None):
    """render `sc` into attachments with given initial contents (loadOp LOAD when the scene has no clears)"""
    col = np.full((sc.height, sc.width, 4), 0xCD, np.uint8) if color0 is None else color0.copy()
    dep = depth0.copy() if depth0 is not None else None
    return scenes.BoundScene(be, sc, color=col, depth=dep).run()


@pytest.mark.parametrize("op", [abi.CMP_LESS, abi.CMP_LEQUAL, abi.CMP_GREATER, abi.CMP_GEQUAL, abi.CMP_EQUAL,
                                abi.CMP_NOTEQUAL, abi.CMP_ALWAYS, abi.CMP_NEVER])
@pytest.mark.parametrize("write", [True, False])
def test_nan_already_in_the_depth_buffer(gpu, vor, op, write):
    """NaN depth left in the attachment (an uncleared image, or an ALWAYS+write draw with degenerate
    vertices): every ordered comparison against it fails and the pixel keeps its contents; NOT_EQUAL and
    ALWAYS pass (rasterizer.cpp:562-576)."""
    sc = scenes.random_triangles(200, 120, 120, 600 + op, depth_op=op, depth_write=write)
    sc.clear_depth = None
    rng = np.random.default_rng(op)
    depth0 = rng.uniform(0.2, 0.8, size=(120, 200)).astype(np.float32)
    depth0[rng.random((120, 200)) < 0.3] = np.float32("nan")
    depth0[5:9, 7:30] = np.float32("-nan")
    a_c, a_d = _render_with(gpu, sc, depth0=depth0)
    b_c, b_d = _render_with(vor, sc, depth0=depth0)
    assert np.array_equal(a_d.view(np.uint32), b_d.view(np.uint32))
    assert np.array_equal(a_c, b_c)


def test_geometry_in_the_last_tile_of_an_8192_target(gpu, vor):
    """tile (255, 255) of an 8192 x 8192 target: its packed tile range is all ones in every byte, which
    must not be mistaken for the dead-triangle marker"""
    W = H = 8192
    px = np.array([[8165, 8165], [8190, 8170], [8170, 8191],      # inside the last tile
                   [8100, 8100], [8191, 8120], [8120, 8191],      # straddles the last 3x3 tiles
                   [10, 10], [40, 12], [12, 40]], dtype=np.float64)
    v = np.zeros((9, 8), np.float32)
    v[:, 0] = (px[:, 0] + 0.5) / W * 2 - 1
    v[:, 1] = -((px[:, 1] + 0.5) / H * 2 - 1)
    v[:, 2], v[:, 3] = 0.5, 1.0
    v[:, 4:8] = np.random.default_rng(1).uniform(0, 1, size=(9, 4))
    from harness import shaders
    pipe = scenes.PipelineDesc(shaders.vs_passthrough(), shaders.fs_color(),
                               [(0, abi.FMT_R32G32B32A32_SFLOAT, 32, 0, 0), (1, abi.FMT_R32G32B32A32_SFLOAT, 32, 16, 0)])
    sc = scenes.Scene("last_tile", W, H, [scenes.Draw(pipe, 9, vbs=[(v, 0)])], depth=False)
    a, _ = scenes.render(gpu, sc)
    b, _ = scenes.render(vor, sc)
    assert (a[8160:, 8160:] != np.array([51, 51, 51, 255], np.uint8)).any()
    assert np.array_equal(a, b)


def test_clear_target_truncation_and_one_byte_targets(gpu, vor):
    """ClearTarget: byte(f * 255.0f) wraps for out-of-range and negative colours exactly as the x86
    truncating convert does (rasterizer.cpp:341-345); a 1-byte-per-pixel target is memset with the red
    channel (:347-350); other pixel sizes are left alone"""
    for col in [(0.2, 0.2, 0.2, 1.0), (0.999, 0.5, 0.0039, 0.25), (1.5, -0.1, 0.7, 2.0), (-3.7, 300.0, 1e12, -1e12),
                (float("nan"), float("inf"), -0.0, 1.0039)]:
        for (w, h) in [(8, 8), (333, 77)]:
            a = np.full((h, w, 4), 9, np.uint8)
            b = np.full((h, w, 4), 9, np.uint8)
            gpu.ClearTarget(abi.make_image(a, w, h, abi.FMT_B8G8R8A8_UNORM), col)
            gpu.flush()
            vor.ClearTarget(abi.make_image(b, w, h, abi.FMT_B8G8R8A8_UNORM), col)
            assert np.array_equal(a, b), col
            a1 = np.full((h, w), 9, np.uint8)
            b1 = np.full((h, w), 9, np.uint8)
            gpu.ClearTarget(abi.make_image(a1, w, h, abi.FMT_R8_UNORM, bpp=1), col)
            gpu.flush()
            vor.ClearTarget(abi.make_image(b1, w, h, abi.FMT_R8_UNORM, bpp=1), col)
            assert np.array_equal(a1, b1), col
    a2 = np.full((8, 8, 2), 9, np.uint8)
    gpu.ClearTarget(abi.make_image(a2, 8, 8, 0, bpp=2), (1.0, 1.0, 1.0, 1.0))
    gpu.flush()
    assert (a2 == 9).all()


def test_results_of_the_same_submit_are_not_overwritten_by_uploads(gpu, vor):
    """coherent (host pointer) memory, no flush between the operations: (a) vkCmdCopyBufferToImage into a
    host-visible image that the next draw samples, (b) an image rendered by one draw and sampled by the
    next, (c) a vertex buffer partly filled by vkCmdCopyBuffer and then read whole by a draw. Each reader
    must see the device-side result, and the flush must bring back exactly what the reference's host
    memory would hold."""
    L = _copy_api(gpu)
    # (a)
    sc = scenes.c2_cube(320, 180)
    want_c, want_d = scenes.render(vor, sc)
    d = sc.draws[0]
    s, b, tex, tw, th, fmt, bpp, layers = d.textures[0]
    staging = np.ascontiguousarray(tex).reshape(-1).copy()
    image = np.zeros(staging.size, np.uint8)             # host-visible, not registered, stale on the host
    im = abi.make_image(image, tw, th, fmt, bpp=bpp, layers=layers)
    sb = abi.make_buffer(staging)
    d.textures = [(s, b, image.reshape(tex.shape), tw, th, fmt, bpp, layers)]
    bound = scenes.BoundScene(gpu, sc)
    gpu.check(L.vb200_copy_buffer_to_image(C.byref(sb), 0, C.byref(im), 0, 0), "copy_buffer_to_image")
    got_c, got_d = bound.run()
    assert np.array_equal(got_c, want_c) and np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
    assert np.array_equal(image, staging)
    # (b)
    def two_pass(be):
        first = scenes.random_triangles(64, 64, 60, 77, has_depth=False, depth_op=abi.CMP_ALWAYS)
        second = scenes.c2_cube(320, 180)
        target = np.full((64, 64, 4), 0xCD, np.uint8)
        dd = second.draws[0]
        s_, b_, _, _, _, fmt_, bpp_, layers_ = dd.textures[0]
        dd.textures = [(s_, b_, target, 64, 64, fmt_, bpp_, layers_)]
        b1 = scenes.BoundScene(be, first, color=target)
        b2 = scenes.BoundScene(be, second)
        b1.submit()
        b2.submit()
        be.flush()
        return target, b2.color, b2.depth
    for x, y in zip(two_pass(gpu), two_pass(vor)):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    # (c)
    full = scenes.random_triangles(200, 120, 80, 78)
    want_c, want_d = scenes.render(vor, full)
    vfull = full.draws[0].vbs[0][0]
    good = vfull.copy()
    half = (good.nbytes // 2) & ~31
    stale = good.copy()
    stale.view(np.uint8).reshape(-1)[:half] = 0xEE        # first half only arrives through the copy
    full.draws[0].vbs = [(stale, 0)]
    stagebuf = good.view(np.uint8).reshape(-1)[:half].copy()
    bound = scenes.BoundScene(gpu, full)
    gpu.check(L.vb200_copy_buffer(C.byref(abi.make_buffer(stagebuf)), 0, C.byref(abi.make_buffer(stale)), 0, half),
              "copy_buffer")
    got_c, got_d = bound.run()
    assert np.array_equal(got_c, want_c) and np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
    assert np.array_equal(stale, good)


def test_mirror_merge_in_the_middle_of_a_frame(gpu, vor):
    """Host ranges are mirrored on first use and mirrors of overlapping ranges are merged (the merged mirror
    lives at a new device address). A draw recorded into the open batch but not launched yet must not be left
    behind writing the old mirror: frame 2's second draw reads a vertex buffer that straddles the end of the
    stale mirror frame 1 left behind, and that stale mirror also holds frame 2's attachments."""
    w, h = 200, 120
    pool = np.zeros(4 << 20, np.uint8)
    img_bytes = w * h * 4
    # frame 1: one large range (as an application's earlier, larger swapchain image would leave behind)
    first = scenes.random_triangles(512, 256, 40, 81, has_depth=False, depth_op=abi.CMP_ALWAYS)
    big_color = pool[:512 * 256 * 4].reshape(256, 512, 4)
    scenes.BoundScene(gpu, first, color=big_color).run()
    stale_end = big_color.nbytes
    # frame 2: attachments inside that range; draw 1 (clears folded in) and draw 2 from different vertex buffers
    a = scenes.random_triangles(w, h, 12, 82, max_size=1.5)
    b = scenes.random_triangles(w, h, 300, 83, max_size=0.05)
    vb_b = b.draws[0].vbs[0][0]
    lo = (stale_end - 64) & ~31                       # straddles the end of frame 1's mirror
    dst = pool[lo:lo + vb_b.nbytes].view(vb_b.dtype).reshape(vb_b.shape)
    dst[...] = vb_b
    want_scene = scenes.random_triangles(w, h, 12, 82, max_size=1.5)
    want_scene.draws += scenes.random_triangles(w, h, 300, 83, max_size=0.05).draws
    want_c, want_d = scenes.render(vor, want_scene)
    b.draws[0].vbs = [(dst, 0)]
    a.draws += b.draws
    color = pool[:img_bytes].reshape(h, w, 4)
    depth = pool[img_bytes:2 * img_bytes].view(np.float32).reshape(h, w)
    color[...] = 0xCD
    depth[...] = 0.75
    got_c, got_d = scenes.BoundScene(gpu, a, color=color, depth=depth).run()
    assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
    assert np.array_equal(got_c, want_c)


def test_mirror_merge_inside_one_draw(gpu, vor):
    """The merge can also happen between two resolves of ONE draw: the colour attachment lies inside the
    stale mirror of an earlier frame, the depth attachment straddles its end. The colour address taken before
    the merge points into the retired allocation; the draw must not render there."""
    w, h = 200, 120
    img_bytes = w * h * 4
    for load in (False, True):
        pool = np.zeros(4 << 20, np.uint8)
        first = scenes.random_triangles(512, 256, 40, 84, has_depth=False, depth_op=abi.CMP_ALWAYS)
        big_color = pool[:512 * 256 * 4].reshape(256, 512, 4)
        scenes.BoundScene(gpu, first, color=big_color).run()
        stale_end = big_color.nbytes
        sc = scenes.random_triangles(w, h, 60, 85)
        want = scenes.random_triangles(w, h, 60, 85)
        if load:
            sc.clear_color = sc.clear_depth = want.clear_color = want.clear_depth = None
        color = pool[:img_bytes].reshape(h, w, 4)
        lo = (stale_end - img_bytes // 2) & ~255
        depth = pool[lo:lo + img_bytes].view(np.float32).reshape(h, w)
        color[...] = 0xCD
        depth[...] = 0.75
        want_c, want_d = scenes.render(vor, want)    # (attachments start as 0xCD / 0.75 there too)
        got_c, got_d = scenes.BoundScene(gpu, sc, color=color, depth=depth).run()
        assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
        assert np.array_equal(got_c, want_c)


def test_clear_fusion_variants(gpu, vor):
    """deferred clears: consumed by the first draw, materialised for a second target/draw, depth clear
    without a depth-using pipeline, colour cleared but depth loaded"""
    a = scenes.random_triangles(300, 200, 100, 90)
    a.clear_depth = None                       # colour cleared, depth loaded from host
    _check(gpu, vor, a)
    b = scenes.random_triangles(300, 200, 100, 91, depth_op=abi.CMP_ALWAYS, depth_write=False)
    _check(gpu, vor, b)                        # depth attachment cleared but never touched by the draw
    c = scenes.random_triangles(333, 211, 100, 92)
    c.draws += scenes.random_triangles(333, 211, 100, 93, blend=(abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, 0),
                                       depth_op=abi.CMP_LESS, depth_write=False).draws
    _check(gpu, vor, c)
    d = scenes.random_triangles(64, 48, 3, 94, max_size=0.05)   # most tiles untouched: they still get cleared
    _check(gpu, vor, d)


def test_draw_edge_cases(gpu, vor):
    """empty and ragged draws: zero vertices, a trailing partial triangle (dropped, rasterizer.cpp:128-149),
    a sub-range selected with `first`, an unsupported topology (draws nothing, :235), and an
    image whose size is not a multiple of the tile"""
    import copy
    base = scenes.random_triangles(97, 65, 40, 31)
    d0 = base.draws[0]
    variants = []
    for count, first in [(0, 0), (1, 0), (2, 0), (4, 0), (7, 3), (d0.count - 3, 3), (d0.count, 0)]:
        sc = copy.deepcopy(base)
        sc.draws[0].count, sc.draws[0].first = count, first
        variants.append(sc)
    sc = copy.deepcopy(base)
    sc.draws[0].pipe.topology = 1    # VK_PRIMITIVE_TOPOLOGY_LINE_LIST: "Unsupported primitive topology!"
    variants.append(sc)
    idx = scenes.random_triangles(97, 65, 40, 32, index_type=abi.INDEX_U16)
    for count, first in [(0, 0), (5, 1), (idx.draws[0].count - 6, 6)]:
        sc = copy.deepcopy(idx)
        sc.draws[0].count, sc.draws[0].first = count, first
        variants.append(sc)
    for sc in variants:
        _check(gpu, vor, sc)


def test_sort_first_ownership_and_exchange_kernels(gpu, vor):
    """tile ownership on ONE GPU: each 'rank' renders only the tiles it owns (tile % world == rank); the
    union of the owned pixels, the pack -> concatenate -> unpack path (k_tiles_pack/_unpack against the
    numpy mirror in harness/tiles.py) and the fused path (peer colour stores, here into buffers on the
    same device) must all reproduce the single-rank image"""
    from harness import tiles
    L = gpu.lib
    L.vb200_set_tile_owner.argtypes = [C.c_int, C.c_int]
    L.vb200_tiles_pack.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64]
    L.vb200_tiles_unpack.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64, C.c_int]
    L.vb200_tiles_per_rank.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.vb200_tiles_per_rank.restype = C.c_uint32
    L.vb200_mem_device_ptr.argtypes = [C.c_void_p]
    L.vb200_mem_device_ptr.restype = C.c_void_p
    L.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mem_unregister.argtypes = [C.c_void_p]
    L.vb200_mem_download.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_set_peer_targets.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int]
    W, H, world = 333, 211, 3
    blend = (abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, 0)
    for kwargs in ({}, {"blend": blend, "depth_write": False}):    # resolve path, ordered path
        sc = scenes.random_triangles(W, H, 150, 41, **kwargs)
        want, _ = scenes.render(vor, sc)
        slots = L.vb200_tiles_per_rank(W, H, world)
        assert slots == tiles.tiles_per_rank(W, H, world)
        gathered = np.zeros((world, slots * 4096), np.uint8)
        union = np.zeros_like(want)
        peers = [np.zeros((H, W, 4), np.uint8) for _ in range(2)]    # stand-ins for two other ranks' images
        try:
            for pbuf in peers + [gathered]:
                gpu.check(L.vb200_mem_register(pbuf.ctypes.data, pbuf.nbytes), "mem_register")
            for r in range(world):
                gpu.check(L.vb200_set_tile_owner(r, world), "set_tile_owner")
                bound = scenes.BoundScene(gpu, sc)
                ptrs = (C.c_void_p * 2)(*[L.vb200_mem_device_ptr(pbuf.ctypes.data) for pbuf in peers])
                dev_color = None
                bound.submit()    # first submit creates the colour mirror; peer targets are keyed by its device address
                gpu.flush()
                dev_color = L.vb200_mem_device_ptr(bound.color.ctypes.data)
                gpu.check(L.vb200_set_peer_targets(dev_color, ptrs, 2), "set_peer_targets")
                bound.submit()
                send = gathered[r]    # registered below: the pack kernel writes device memory
                gpu.check(L.vb200_tiles_pack(C.byref(bound.color_img), L.vb200_mem_device_ptr(send.ctypes.data),
                                             send.nbytes), "tiles_pack")
                gpu.check(L.vb200_mem_download(send.ctypes.data, send.nbytes), "mem_download")
                gpu.flush()
                gpu.check(L.vb200_set_peer_targets(dev_color, None, 0), "set_peer_targets")
                own = tiles.owned_mask(W, H, r, world)
                assert np.array_equal(bound.color[own], want[own]), f"rank {r}: owned pixels differ"
                union[own] = bound.color[own]
                assert np.array_equal(send, tiles.pack(np.where(own[..., None], bound.color, 0).astype(np.uint8), r, world))
            gpu.check(L.vb200_set_tile_owner(0, 1), "set_tile_owner")
            assert np.array_equal(union, want)
            # un-tile the "all-gathered" buffer on the device
            out = np.zeros((H, W, 4), np.uint8)
            im = abi.make_image(out, W, H, abi.FMT_B8G8R8A8_UNORM)
            flat = gathered.reshape(-1)    # already resident in its mirror (the pack kernels wrote it there)
            gpu.check(L.vb200_tiles_unpack(C.byref(im), L.vb200_mem_device_ptr(flat.ctypes.data), flat.nbytes, world),
                      "tiles_unpack")
            gpu.flush()
            assert np.array_equal(out, want)
            assert np.array_equal(out, tiles.unpack(flat, W, H, world))
            # fused exchange: every rank stored its pixels into both "peer" images
            for pbuf in peers:
                gpu.check(L.vb200_mem_download(pbuf.ctypes.data, pbuf.nbytes), "mem_download")
            gpu.flush()
            for pbuf in peers:
                assert np.array_equal(pbuf, want)
        finally:
            L.vb200_set_tile_owner(0, 1)
            for pbuf in peers + [gathered]:
                L.vb200_mem_unregister(pbuf.ctypes.data)


def test_present_is_ordered_against_later_frames(gpu, vor):
    """vb200_present copies on a second stream; a later frame that overwrites the SAME image (a
    single-buffered application) must not disturb the copy, and a double-buffered one gets both frames"""
    L = gpu.lib
    L.vb200_present.argtypes = [C.POINTER(abi.Image), C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    L.vb200_present_wait.argtypes = [C.c_int]
    L.vb200_mem_register.argtypes = [C.c_void_p, C.c_uint64]
    L.vb200_mem_unregister.argtypes = [C.c_void_p]
    W, H = 1024, 768    # large enough for the copy to still be in flight when the next frame starts
    sc1 = scenes.random_triangles(W, H, 400, 51)
    sc2 = scenes.random_triangles(W, H, 400, 52)
    want1, _ = scenes.render(vor, sc1)
    want2, _ = scenes.render(vor, sc2)
    shown = [np.zeros((H, W, 4), np.uint8) for _ in range(2)]
    try:
        for a in shown:
            gpu.check(L.vb200_mem_register(a.ctypes.data, a.nbytes), "mem_register")
        b1 = scenes.BoundScene(gpu, sc1)
        b2 = scenes.BoundScene(gpu, sc2, color=b1.color, depth=b1.depth)    # same attachments
        t1, t2 = C.c_int(), C.c_int()
        b1.submit()
        gpu.check(L.vb200_present(C.byref(b1.color_img), shown[0].ctypes.data, shown[0].nbytes, C.byref(t1)), "present")
        b2.submit()                                                          # overwrites the image being copied
        gpu.check(L.vb200_present(C.byref(b2.color_img), shown[1].ctypes.data, shown[1].nbytes, C.byref(t2)), "present")
        gpu.check(L.vb200_present_wait(t1.value), "present_wait")
        assert np.array_equal(shown[0], want1)
        gpu.check(L.vb200_present_wait(t2.value), "present_wait")
        assert np.array_equal(shown[1], want2)
        gpu.flush()
        assert np.array_equal(b1.color, want2)                               # coherent mode still downloads
        assert L.vb200_present(C.byref(b1.color_img), shown[0].ctypes.data, 16, C.byref(t1)) != 0    # too small
    finally:
        gpu.flush()
        for a in shown:
            L.vb200_mem_unregister(a.ctypes.data)
