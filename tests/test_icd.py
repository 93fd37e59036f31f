"""Drop-in at the Vulkan level: the same headless Vulkan application (tests/icd/vk_driver.cpp:
vkCreateInstance .. vkCmdDrawIndexed .. vkQueueSubmit, attachments read back through vkMapMemory) runs
against
  * the reference ICD  — every reference .cpp unmodified (oracle/_ref/libvisor_ref.so), and
  * the CUDA ICD       — the reference's ICD surface (icd_interface/cmd_record/shaders/images/... unmodified)
                         with rasterizer.cpp, texture_sampling.cpp, spirv_compile.cpp, cmd_exec.cpp, memory.cpp
                         replaced by integration/*.cpp over the C-ABI (integration/libvisor_b200_icd.so).
"""
import numpy as np
import pytest

from harness import abi, scenes, vkdriver

needs_driver = pytest.mark.skipif(not vkdriver.available(), reason="oracle/_ref/libvk_driver.so not built")

SCENES = {
    "c1": lambda: scenes.c1_triangle(640, 360),
    "c2": lambda: scenes.c2_cube(640, 360),
    "c3": lambda: scenes.c3_mesh(640, 360, 160, 80),
    "c4": lambda: scenes.c4_particles(640, 360, 8000),
    "c5": lambda: scenes.c5_textured(640, 360, 160, 80, tex_size=128),
    "strip_u16": lambda: scenes.random_triangles(320, 200, 62, 9, topology=abi.TOPO_STRIP, index_type=abi.INDEX_U16),
    "load_op_load": lambda: _no_clear(scenes.random_triangles(200, 120, 50, 3)),
    "c2_bc2": lambda: _bc_cube(abi.FMT_BC2_UNORM_BLOCK),
    "c2_bc3": lambda: _bc_cube(abi.FMT_BC3_UNORM_BLOCK),
}


def _bc_cube(fmt):
    """the C2 cube sampling a block-compressed texture uploaded through a staging buffer +
    vkCmdCopyBufferToImage into DEVICE_LOCAL memory"""
    sc = scenes.c2_cube(320, 180)
    d = sc.draws[0]
    s, b, _, _, _, _, _, layers = d.textures[0]
    d.textures = [(s, b, scenes.bc_blocks(np.random.default_rng(21), 128, 64), 128, 64, fmt, 1, layers)]
    return sc


def _no_clear(sc):
    sc.clear_color = None
    sc.clear_depth = None
    return sc


@needs_driver
@pytest.mark.parametrize("name", sorted(SCENES))
def test_reference_icd_equals_operator_level_oracle(vor, name):
    """The Vulkan call sequence through the reference ICD gives the image the operator-level oracle gives:
    validates the driver and that the ICD layers add nothing to the result."""
    sc = SCENES[name]()
    c_icd, d_icd, _ = vkdriver.run(vkdriver.ICD_REF, sc)
    c, d = scenes.render(vor, sc)
    assert np.array_equal(c_icd, c)
    if d is not None:
        assert np.array_equal(d_icd.view(np.uint32), d.view(np.uint32))


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_cuda_icd_equals_reference_icd(name):
    sc = SCENES[name]()
    c_ref, d_ref, _ = vkdriver.run(vkdriver.ICD_REF, sc)
    c_gpu, d_gpu, _ = vkdriver.run(vkdriver.ICD_CUDA, sc)
    assert np.array_equal(c_gpu, c_ref), f"{name}: {(c_gpu != c_ref).any(-1).sum()} pixels differ"
    if d_ref is not None:
        assert np.array_equal(d_gpu.view(np.uint32), d_ref.view(np.uint32))
