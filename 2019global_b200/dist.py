"""Framebuffer gather to rank 0: one process per GPU, torch.distributed plumbing.

The path shards by image tile with no data-path collective; the only exchange is this gather.
Every rank contributes its compact tile array (1/world of the frame), padded to the largest
rank's length so that torch.distributed.gather (grouped ncclSend/ncclRecv over NVLink) can be
used; rank 0 scatters each payload into the frame with `untile`.
"""
import torch
import torch.distributed as dist

from . import tiles


def gather_frame(local, w, h, channels, frame, untile, group=None):
    """local: this rank's compact array, flat, padded to padded_len(w,h,world)*channels.
    frame: rank 0's full frame buffer (ignored elsewhere). untile(rank, payload, frame) scatters."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        untile(0, local, frame)
        return frame
    if rank == 0:
        bufs = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, bufs, dst=0, group=group)
        for r in range(world):
            untile(r, bufs[r], frame)
        return frame
    dist.gather(local, None, dst=0, group=group)
    return None


def padded_len(w, h, world):
    """Compact-array length every rank pads to (rank 0 owns the most tiles)."""
    return tiles.n_local_tiles(w, h, 0, world) * tiles.TILE_PIX


def shared_frame(rt, w, h, group=None):
    """One SharedFrame per job: rank 0 allocates it, the CUDA IPC blob travels by broadcast, every
    other rank maps it. After this the data path needs no collective: each rank's resolve kernel
    stores its pixels into rank 0's HBM over NVLink (engine.SharedFrame)."""
    from .engine import SharedFrame
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if rank == 0:
        f = SharedFrame.create(rt, w, h)
        box = [f.blob()]
    else:
        f, box = None, [None]
    if world > 1:
        dist.broadcast_object_list(box, src=0, group=group)
        if rank != 0:
            f = SharedFrame.attach(rt, box[0], w, h)
    return f
