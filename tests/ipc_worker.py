"""One rank of tests/test_gpu_parity.py::test_two_processes_share_queue_and_frame_over_ipc (one process per GPU).

    python tests/ipc_worker.py RANK WORLD PORT OBJ_PATH TMP_DIR

The 128 bytes of IPC handles travel over a gloo process group (host memory); the data path is the kernel itself.
"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    rank, world, port, obj = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import ripoff_raytracer_b200 as rr
    from ripoff_raytracer_b200 import _abi, multigpu

    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, spp, bounces = 640, 360, 4, 12
    cam = rr.default_camera(W, H)
    r = rr.Renderer((rank,))
    r.upload(rr.default_scene(obj))
    q, f = multigpu.exchange_handles(dist, rank, lambda: r.queue_export(W, H))
    if rank != 0:
        r.queue_import(W, H, q, f)
    assert multigpu.negotiate_mode(dist, True) == "shared"
    tiles = rays = 0
    for frame in range(3):  # several frames: the reset / barrier protocol of include/rr_api.h, epochs 1, 2, 3
        dist.barrier()
        if rank == 0:
            r.queue_reset()
        dist.barrier()
        st = r.render_shared(cam, W, H, spp, bounces)
        tiles, rays = st["tiles"], st["rays"]
    dist.barrier()
    import torch

    tot = torch.tensor([tiles, rays], dtype=torch.int64)
    dist.all_reduce(tot)
    if rank == 0:
        got = r.read_frame(W, H)
        one = rr.Renderer((0,))
        one.upload(rr.default_scene(obj))
        want, _, st1 = one.render(cam, W, H, spp, bounces)
        one.close()
        print(f"frame_equal={bool(np.array_equal(got, want))}")
        print(f"tiles_ok={int(tot[0]) == st1['tiles'] and int(tot[1]) == st1['rays']} ({int(tot[0])} tiles, {int(tot[1])} rays)")
    # misuse is reported: (a) a frame larger than the exported one, (b) a rank rendering without the reset
    # (the counter still holds the finished frame's epoch)
    reported = 0
    try:
        r.render_shared(rr.default_camera(W * 2, H), W * 2, H, 1, 1)
    except _abi.RRError as e:
        reported += e.status == 10
    dist.barrier()
    try:
        r.render_shared(cam, W, H, 1, 1)  # epoch 4 expected, the counter says 3
    except _abi.RRError as e:
        reported += e.status == 10
    if rank == 0:
        print(f"misuse_reported={reported == 2}")
    else:
        assert reported == 2
    dist.barrier()
    r.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
