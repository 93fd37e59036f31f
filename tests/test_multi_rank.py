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


def _io_worker(rank: int, world: int, port: int, q) -> None:
    """the N-rank end-to-end data movement of bench.py with numpy standing in for HBM: every rank
    'uploads' its slice of each input, the slices are all-gathered in place, and each rank writes its
    band of rows of the finished image into one shared buffer"""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from multiprocessing import shared_memory

    from harness import tiles
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ok = True
    rng = np.random.default_rng(7)
    for nbytes in (1 << 20, (1 << 20) + 12345, 3 * (1 << 18) + 7, 1000):
        host = rng.integers(0, 256, size=nbytes, dtype=np.uint8)    # same on every rank (same seed)
        mirror = np.zeros(nbytes, np.uint8)                          # this rank's "HBM mirror"
        sl, tail = tiles.upload_shard(nbytes, world)
        if sl == 0:
            mirror[:] = host
        else:
            mirror[rank * sl:(rank + 1) * sl] = host[rank * sl:(rank + 1) * sl]
            if tail:
                mirror[sl * world:] = host[sl * world:]
            full = torch.from_numpy(mirror[:sl * world])
            dist.all_gather_into_tensor(full, full[rank * sl:(rank + 1) * sl].clone())
        ok &= bool(np.array_equal(mirror, host))
    # band download into one shared host buffer
    W, H = 333, 97
    image = np.arange(W * H, dtype=np.int32).reshape(H, W)    # what every rank holds after the exchange
    name = f"vb200_test_{port}"
    shm = shared_memory.SharedMemory(name=name, create=True, size=W * H * 4) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        shm = shared_memory.SharedMemory(name=name)
        try:
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    frame = np.ndarray((H, W), dtype=np.int32, buffer=shm.buf)
    lo, hi = tiles.row_band(H, rank, world)
    frame[lo:hi] = image[lo:hi]
    dist.barrier()
    if rank == 0:
        ok &= bool(np.array_equal(frame, image))
    del frame
    gathered = [None] * world
    dist.all_gather_object(gathered, ok)
    dist.barrier()
    shm.close()
    if rank == 0:
        shm.unlink()
        q.put(all(gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_upload_and_band_download_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_io_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_arithmetic():
    from harness import tiles
    for nbytes in (1000, 65536 * 2, 128_873_360, 48_000_000, 16_777_216 + 5):
        for world in (2, 4, 8):
            sl, tail = tiles.upload_shard(nbytes, world)
            assert sl % 256 == 0 and sl * world + tail == nbytes and 0 <= tail
            assert sl == 0 or tail < world * 256 + world
    for h, world in ((4320, 8), (97, 4), (5, 8)):
        bands = [tiles.row_band(h, r, world) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == h
        assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
