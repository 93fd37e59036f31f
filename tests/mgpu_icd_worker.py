"""One rank of the sort-first ICD test: a headless Vulkan application (tests/icd/vk_driver.cpp) on the CUDA ICD,
with the group described by RANK / WORLD_SIZE / VISOR_B200_SESSION in the environment (integration/gpu_b200.cpp)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harness import abi, scenes, vkdriver  # noqa: E402


def main():
    rank = int(sys.argv[1])
    cpu = abi.backend("vor")
    for sc in (scenes.c3_mesh(640, 360, 160, 80), scenes.c2_cube(320, 180), scenes.c4_particles(320, 200, 3000)):
        want_c, _ = scenes.render(cpu, sc)
        got_c, _, _ = vkdriver.run(vkdriver.ICD_CUDA, sc)
        assert np.array_equal(got_c, want_c), f"rank {rank} {sc.name}: {(got_c != want_c).any(-1).sum()} pixels differ"
    print(f"rank {rank}: ok", flush=True)


if __name__ == "__main__":
    main()
