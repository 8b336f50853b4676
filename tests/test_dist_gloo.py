"""N>1 host logic on CPU: tile partition + framebuffer gather over torch.distributed (gloo, world 2)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

W, H = 173, 90  # ragged: partial tiles on both edges


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = importlib.import_module("2019global_b200.tiles")
    g19dist = importlib.import_module("2019global_b200.dist")
    g = tiles.local_pixels(W, H, rank, world)
    pad = g19dist.padded_len(W, H, world)
    # the "render": each owned pixel carries (global index, rank, 7); padding carries garbage
    local = torch.full((pad * 3,), -5.0)
    vals = np.stack([g.astype(np.float32), np.full(g.shape, rank, np.float32), np.full(g.shape, 7, np.float32)], 1)
    local[: g.shape[0] * 3] = torch.from_numpy(vals.reshape(-1))
    frame = torch.full((H * W * 3,), -1.0) if rank == 0 else None

    def untile(r, payload, fr):
        f = fr.numpy().reshape(H * W, 3)
        tiles.untile_numpy(f, payload.numpy().reshape(-1, 3), W, H, r, world)

    res = g19dist.gather_frame(local, W, H, 3, frame, untile)
    if rank == 0:
        np.save(out_path, res.numpy().reshape(H * W, 3))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_reassembles_the_frame(tmp_path, world):
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    f = np.load(out)
    assert np.array_equal(f[:, 0], np.arange(H * W, dtype=np.float32))  # every pixel, once, in place
    assert (f[:, 2] == 7).all()
    tiles = importlib.import_module("2019global_b200.tiles")
    tx = (W + 31) // 32
    x, y = np.arange(H * W) % W, np.arange(H * W) // W
    assert np.array_equal(f[:, 1], ((y // 32) * tx + x // 32) % world)  # owner = tile % world


def test_single_process_gather_is_identity():
    tiles = importlib.import_module("2019global_b200.tiles")
    g19dist = importlib.import_module("2019global_b200.dist")
    g = tiles.local_pixels(W, H, 0, 1)
    local = torch.from_numpy(np.repeat(g.astype(np.float32), 3))
    frame = torch.zeros(H * W * 3)

    def untile(r, payload, fr):
        tiles.untile_numpy(fr.numpy().reshape(H * W, 3), payload.numpy().reshape(-1, 3), W, H, r, 1)

    g19dist.gather_frame(local, W, H, 3, frame, untile)
    assert np.array_equal(frame.numpy().reshape(-1, 3)[:, 0], np.arange(H * W, dtype=np.float32))
