"""World-size-2 checks of the multi-GPU host logic on CPU (gloo): handle exchange, mode negotiation, the static
tile partition and its one-reduce gather (ripoff_raytracer_b200/multigpu.py).  The device side of the shared
atomic tile queue is covered by bench.py --gpus N on real GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ripoff_raytracer_b200 import multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, fail_rank, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. rank 0 "exports" two 64-byte handles, everyone receives them
        q0 = np.arange(64, dtype=np.uint8)
        f0 = (np.arange(64, dtype=np.uint8) * 3) % 251
        q, f = multigpu.exchange_handles(dist, rank, lambda: (q0, f0))
        assert np.array_equal(q, q0) and np.array_equal(f, f0)
        # 2. one rank failing to attach sends EVERY rank to the static partition
        mode = multigpu.negotiate_mode(dist, attached_ok=(rank != fail_rank))
        assert mode == ("strided" if fail_rank is not None else "shared")
        # 3. static partition: each rank paints its tiles with (tile id + 1), rank 0 reduces
        tx, ty = multigpu.tile_grid(W, H)
        frame = np.zeros((H, W, 4), np.uint8)
        mine = multigpu.strided_tiles(rank, world, tx * ty)
        for t in mine:
            x0, y0, w, h = multigpu.tile_rect(int(t), W, H)
            frame[y0:y0 + h, x0:x0 + w, :3] = (int(t) + 1) % 251
            frame[y0:y0 + h, x0:x0 + w, 3] = 255
        merged = multigpu.merge_strided_frames(dist, frame, rank)
        if rank == 0:
            want = np.zeros((H, W, 4), np.uint8)
            for t in range(tx * ty):
                x0, y0, w, h = multigpu.tile_rect(t, W, H)
                want[y0:y0 + h, x0:x0 + w, :3] = (t + 1) % 251
                want[y0:y0 + h, x0:x0 + w, 3] = 255
            assert np.array_equal(merged, want)
            out.put("ok")
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [None, 1])
def test_world_size_2_host_logic(fail_rank):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    W, H = 37, 18  # ragged against the 8x4 tile
    procs = [ctx.Process(target=_worker, args=(r, 2, port, W, H, fail_rank, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_tile_partition_covers_every_tile_once():
    for world in (1, 2, 3, 8):
        n = 1000
        seen = np.concatenate([multigpu.strided_tiles(r, world, n) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(n))
    with pytest.raises(ValueError):
        multigpu.strided_tiles(2, 2, 10)
    assert multigpu.tile_grid(3840, 2160) == (480, 540)
    assert multigpu.tile_rect(479, 3840, 2160) == (3832, 0, 8, 4)
    assert multigpu.tile_rect(4, 37, 18) == (32, 0, 5, 4)


def test_queue_items_cover_every_pixel_once():
    """The queue counter hands out pixels numbered tile by tile (include/rr_api.h, csrc/rr_render.cu pixel phase): over all
    ranks of a static partition, and for the shared queue (world 1), every pixel of the frame appears exactly once and the
    padding items of ragged border tiles map to no pixel."""
    for (W, H), (tw, th) in [((37, 23), (8, 4)), ((8, 4), (8, 4)), ((9, 5), (5, 5)), ((1, 130), (32, 32)), ((64, 32), (8, 4))]:
        for world in (1, 3):
            seen = np.zeros((H, W), np.int32)
            padding = 0
            for rank in range(world):
                n = multigpu.queue_items(W, H, tw, th, rank, world)
                assert n % (tw * th) == 0
                for item in range(n):
                    px = multigpu.item_pixel(item, W, H, tw, th, rank, world)
                    if px is None:
                        padding += 1
                    else:
                        seen[px[1], px[0]] += 1
            assert np.all(seen == 1), (W, H, tw, th, world)
            tx, ty = multigpu.tile_grid(W, H, tw, th)
            assert padding == tx * ty * tw * th - W * H
