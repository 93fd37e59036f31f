"""Sort-first over several GPUs through the C-ABI alone (vb200_mgpu_*): one process per GPU, bootstrap over
unix sockets, symmetric buffers with NVSwitch multicast (or peer stores), device-side barrier, sliced input
upload. No torch / NCCL anywhere on the data path. Needs >= 2 GPUs (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_mgpu.py -m gpu`)."""
import os
import subprocess
import sys
import uuid

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30)
        return sum(1 for l in out.stdout.splitlines() if l.startswith("GPU ")) if out.returncode == 0 else 0
    except Exception:
        return 0


def _run_group(world, mode, script="mgpu_worker.py", extra_env=None):
    session = uuid.uuid4().hex[:12]
    procs = []
    for rank in range(world):
        env = dict(os.environ)
        env.update(extra_env(rank, world, session) if extra_env else {})
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", script), str(rank), str(world),
                                       session, mode], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                      env=env))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out[-3000:]}"
        assert f"rank {rank}: ok" in out
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["alloc", "alloc-p2p", "mirrors", "mirrors-p2p"])
def test_two_ranks_render_the_whole_image(mode):
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    _run_group(2, mode)


@pytest.mark.gpu
def test_all_gpus_of_the_box():
    n = _gpu_count()
    if n < 3:
        pytest.skip("needs more than two GPUs")
    _run_group(min(n, 8), "alloc")


@pytest.mark.gpu
def test_cuda_icd_renders_sort_first_on_two_gpus():
    """the same Vulkan application started twice (RANK/WORLD_SIZE/VISOR_B200_SESSION in the environment): each
    process renders its tiles on its GPU and both end up with the reference image in their framebuffer memory"""
    from harness import vkdriver
    if _gpu_count() < 2 or not vkdriver.available():
        pytest.skip("needs two GPUs and the Vulkan driver harness")
    _run_group(2, "icd", script="mgpu_icd_worker.py",
               extra_env=lambda rank, world, session: {"RANK": str(rank), "WORLD_SIZE": str(world),
                                                       "LOCAL_RANK": str(rank), "VISOR_B200_SESSION": session})
