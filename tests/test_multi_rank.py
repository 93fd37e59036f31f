"""Sort-first host logic on CPU with torch.distributed/gloo, world_size 2 and 4: every rank keeps
only the tiles it owns, packs them owner-major, all-gathers, un-tiles — the assembled image must be
the single-rank image. (The GPU run does the same with NCCL and the CUDA pack/unpack kernels.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, width: int, height: int, q) -> None:
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from harness import abi, scenes, tiles
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    vor = abi.backend("vor")
    sc = scenes.random_triangles(width, height, 120, 21)
    full, _ = scenes.render(vor, sc)
    # what this rank would hold after rendering only its own tiles
    mine = np.where(tiles.owned_mask(width, height, rank, world)[..., None], full, 0).astype(np.uint8)
    send = torch.from_numpy(tiles.pack(mine, rank, world).copy())
    recv = torch.empty(world * send.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, send)
    out = tiles.unpack(recv.numpy(), width, height, world)
    ok = bool(np.array_equal(out, full))
    gathered = [None] * world
    dist.all_gather_object(gathered, ok)
    if rank == 0:
        q.put(all(gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,width,height", [(2, 200, 120), (4, 333, 97)])
def test_sort_first_assemble_gloo(world, width, height):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, width, height, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_tiles_per_rank_matches_library():
    import ctypes as C

    import visor_b200
    from harness import tiles
    L = visor_b200.lib()
    L.vb200_tiles_per_rank.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.vb200_tiles_per_rank.restype = C.c_uint32
    for (w, h, n) in [(7680, 4320, 8), (7680, 4320, 2), (1920, 1080, 4), (333, 97, 4), (32, 32, 8)]:
        assert L.vb200_tiles_per_rank(w, h, n) == tiles.tiles_per_rank(w, h, n)
