"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): coverage, depth-test outcomes and integer/UNORM framebuffer writes
bit-exact; shaded colour within 1 UNORM8 LSB — in practice every scene here is required to be
byte-identical because the shader arithmetic is IEEE-exact in both (no sin/cos/pow in these scenes).
"""
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
    if d_cpu is not None:
        bad = (d_gpu.view(np.uint32) != d_cpu.view(np.uint32)).sum()
        assert bad == 0, f"{sc.name}: {bad} depth words differ"
    diff = np.abs(c_gpu.astype(np.int16) - c_cpu.astype(np.int16))
    if exact:
        assert diff.max() == 0, f"{sc.name}: {(diff > 0).any(-1).sum()} pixels differ (max {diff.max()} LSB)"
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
