"""Host-side plumbing of the multi-GPU tile queue (one process per GPU).

Replaces the reference's `multiThreadedCompute` (src/image.hpp:280-350: one std::thread per device popping
`(tileX, tileY)` from a mutex-guarded std::queue, merging tiles into a shared `pixels` under a second mutex).
Here rank 0 owns ONE 64-bit queue counter and the frame in its HBM; the other ranks attach to both with CUDA IPC
handles and their persistent warps take pixels (numbered tile by tile; a warp adds as many as it has free path
slots) with system-scope atomics and store them straight into rank 0's frame over NVLink (csrc/rr_render.cu pixel
phase, csrc/rr_api.cu rr_queue_*).  What is left for the host is: ship
128 bytes of handles, agree on the mode, and -- only when peer access is unavailable -- a static partition plus
one reduce.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) carries those few bytes; no collective
touches the data path in "shared" mode.
"""
from __future__ import annotations

import numpy as np

HANDLE_BYTES = 64


def tile_grid(width: int, height: int, tile_w: int = 8, tile_h: int = 4):
    """Tiles are numbered row-major (csrc/rr_render.cu: tile % tiles_x, tile / tiles_x)."""
    tx = (width + tile_w - 1) // tile_w
    ty = (height + tile_h - 1) // tile_h
    return tx, ty


def strided_tiles(rank: int, world: int, n_tiles: int) -> np.ndarray:
    """Static partition used when the shared queue cannot be attached: rank r renders tiles r, r+world, ..."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("rank/world")
    return np.arange(rank, n_tiles, world, dtype=np.int64)


def tile_rect(tile: int, width: int, height: int, tile_w: int = 8, tile_h: int = 4):
    tx, _ = tile_grid(width, height, tile_w, tile_h)
    x0, y0 = (tile % tx) * tile_w, (tile // tx) * tile_h
    return x0, y0, min(tile_w, width - x0), min(tile_h, height - y0)


def queue_items(width: int, height: int, tile_w: int = 8, tile_h: int = 4, rank: int = 0, world: int = 1) -> int:
    """Work items the queue counter of one launch runs through (csrc/rr_api.cu fill_params / render_frame): the pixels of
    its tiles, ragged border tiles padded to tile_w x tile_h (the padding items are skipped by the kernel)."""
    tx, ty = tile_grid(width, height, tile_w, tile_h)
    return len(strided_tiles(rank, world, tx * ty)) * tile_w * tile_h


def item_pixel(item: int, width: int, height: int, tile_w: int = 8, tile_h: int = 4, rank: int = 0, world: int = 1):
    """Item -> (x, y) of the pixel the kernel renders for it, or None for a padding item (csrc/rr_render.cu, pixel phase):
    items are numbered tile by tile (the launch's tiles in ascending order), row-major inside a tile."""
    tx, _ = tile_grid(width, height, tile_w, tile_h)
    seq, k = divmod(item, tile_w * tile_h)
    tile = rank + seq * world
    x, y = (tile % tx) * tile_w + k % tile_w, (tile // tx) * tile_h + k // tile_w
    return (x, y) if x < width and y < height else None


def exchange_handles(dist, rank: int, export_fn, device=None):
    """Rank 0 exports (queue handle, frame handle); every rank gets both.  128 bytes over the process group."""
    import torch

    buf = torch.zeros(2 * HANDLE_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        q, f = export_fn()
        buf.copy_(torch.from_numpy(np.concatenate([np.asarray(q, np.uint8), np.asarray(f, np.uint8)])))
    dist.broadcast(buf, 0)
    h = buf.cpu().numpy()
    return h[:HANDLE_BYTES].copy(), h[HANDLE_BYTES:].copy()


def negotiate_mode(dist, attached_ok: bool, device=None) -> str:
    """"shared" only if EVERY rank attached to rank 0's queue and frame; otherwise all ranks fall back together."""
    import torch

    ok = torch.tensor([1.0 if attached_ok else 0.0], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return "shared" if ok.item() > 0 else "strided"


def merge_strided_frames(dist, frame_u8: np.ndarray, rank: int, device=None):
    """Fallback gather: every rank rendered a disjoint tile set into a zeroed frame; one SUM-reduce to rank 0."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(frame_u8).view(np.int32).copy())
    if device is not None:
        t = t.to(device)
    dist.reduce(t, 0, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    return t.cpu().numpy().view(np.uint8).reshape(frame_u8.shape)
