#!/usr/bin/env python
"""bench.py -- throughput of the render hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus 1 --steps K ...  # the reference's CPU implementation of the path

A *step* is one frame of the workload (default: BASELINE.json configs[3], the ~1 M-triangle mixed
mesh+sphere scene at 3840x2160, 256 spp -- the configuration the north-star target is quoted on).
`value` is whole-job Mrays/s with the scene resident in HBM (path segments per second; a segment is
one closest-hit query, reference src/Trace.cl:494).  `e2e` is the same metric through the C-ABI
boundary call with HOST buffers: scene upload from pinned memory + LBVH build + render + frame
read-back, every step.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


class stdout_to_stderr:
    """NCCL prints its version banner / debug lines to fd 1 when the communicator is created; stdout must carry
    exactly one JSON line, so fd 1 points at stderr while the process group and its first collective come up."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


F_BOX, F_TRI, F_SPH = 24, 53, 24     # algorithmic flop per test (sphere: DESIGN.md 6)
B_BOX, B_TRI, B_SPH = 32, 48, 16     # algorithmic bytes per test


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--bounces", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_workload(args):
    from ripoff_raytracer_b200 import workloads

    kw = {}
    for k in ("width", "height", "spp", "bounces"):
        if getattr(args, k):
            kw[k] = getattr(args, k)
    return workloads.WORKLOADS[args.workload](**kw)


# ----------------------------------------------------------------------------- clocks ----
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.25)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) > 2 + k and r[2 + k] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------- reference arm ----
def cpu_reference_run(wl, steps, warmup, budget_s=20.0):
    """The reference's own kernel source (src/Trace.cl compiled for the host, -O3 -ffast-math: the analogue of
    its -cl-fast-relaxed-math JIT flags) with its own SAH hierarchy, on all host cores, on a bounded sample
    of the workload: the same scene and camera at reduced resolution / spp."""
    from oracle.pyoracle import Oracle, Reference

    cores = os.cpu_count() or 1
    t, m, r, sp = wl.scene.arrays()
    # the reference kernel has no sphere primitive: a spheres-only workload can only be timed with the port
    kind = "reference" if Reference.available("fast") and len(t) > 14 else "port"
    # sample geometry: same aspect, 1/8 linear resolution (>= 64 px wide), spp scaled to the time budget
    W = max(64, wl.width // 8)
    H = max(36, wl.height // 8)
    cam = wl.cam.copy()
    if kind == "reference":
        # the reference kernel has no sphere primitive: its arm renders the triangle part of the scene
        ref = Reference("fast")
        t0 = time.perf_counter()
        rt, rm, rn = ref.scene_from_arrays(t, m, r)
        build_s = time.perf_counter() - t0
        counter = Oracle(rt, rm, r)                      # strict restatement: counts the segments of the same sample
        render = lambda spp: ref.render(cam, W, H, spp, wl.bounces, threads=cores)
        sah_counter = Oracle(rt, rm, r, ref_gpunodes=rn)  # the same restatement walking the REFERENCE's own SAH nodes
    else:
        sah_counter = None
        t0 = time.perf_counter()
        counter = Oracle(t, m, r, sp)
        build_s = time.perf_counter() - t0
        render = lambda spp: counter.render(cam, W, H, spp, wl.bounces, threads=cores)
    # calibrate spp for ~budget_s/(steps+warmup) per step
    t0 = time.perf_counter()
    render(1)
    per_spp = max(time.perf_counter() - t0, 1e-4)
    spp = int(max(1, min(wl.spp, budget_s / max(steps + warmup, 1) / per_spp)))
    _, _, cst = counter.render(cam, W, H, spp, wl.bounces, threads=cores)
    rays = cst["rays"]
    for _ in range(warmup):
        render(spp)
    t0 = time.perf_counter()
    for _ in range(steps):
        render(spp)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    sah = None
    if sah_counter is not None:  # test counts of the reference's own hierarchy on the same rays (1 spp: the counts are per ray)
        _, _, s1 = sah_counter.render(cam, W, H, 1, wl.bounces, threads=cores)
        _, _, l1 = counter.render(cam, W, H, 1, wl.bounces, threads=cores)
        sah = {"box_tests_per_ray": s1["box_tests"] / max(s1["rays"], 1), "tri_tests_per_ray": s1["tri_tests"] / max(s1["rays"], 1),
               "lbvh_binary_walk_box_tests_per_ray": l1["box_tests"] / max(l1["rays"], 1),
               "lbvh_binary_walk_tri_tests_per_ray": l1["tri_tests"] / max(l1["rays"], 1),
               "what": "oracle restatement of src/Trace.cl:319-397 on the node list the reference's SplitBVH (src/readobj.hpp:206-267) "
                       "built for this scene, triangles only, same rays; beside it the oracle's binary in-order walk of our LBVH"}
    return {
        "reference_sah": sah,
        "mrays_s": rays / dt / 1e6, "msamples_s": W * H * spp / dt / 1e6, "ms_per_step": dt * 1e3, "cores": cores,
        "kind": kind, "build_s": build_s,
        "sample": f"{wl.name} scene ({len(t)} tris{'' if kind == 'port' else ', triangles only: the reference kernel has no spheres'}) "
                  f"at {W}x{H}, {spp} spp, {wl.bounces} bounces = {rays} path segments per step; "
                  + ("reference Trace.cl text compiled for the host (-O3 -ffast-math) with its SAH BVH" if kind == "reference"
                     else "oracle/rr_oracle.c restatement") + f", {cores} threads",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args)
    res = cpu_reference_run(wl, args.steps, max(args.warmup, 1), budget_s=60.0)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": res["mrays_s"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, args.gpus),
        "msamples_per_s": res["msamples_s"],
        "cpu_baseline": {"value": res["mrays_s"], "unit": "Mrays/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"],
                         "reference_sah": res["reference_sah"]},
        "e2e": {"value": res["mrays_s"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, n_gpus):
    return {"workload": f"{wl.name}: {wl.description}; {wl.width}x{wl.height}, {wl.spp} spp, {wl.bounces} bounces",
            "triangles": wl.scene.n_triangles, "spheres": wl.scene.n_spheres, "meshes": wl.scene.n_meshes,
            "width": wl.width, "height": wl.height, "spp": wl.spp, "max_bounces": wl.bounces,
            "parallelism": f"tile-queue x{n_gpus}" if n_gpus > 1 else "1 GPU, persistent warps over an atomic tile queue",
            "l2": "scene arrays (nodes+triangles+normals) exceed the 126 MB L2 and a 256 MB buffer is written between steps"}


# ------------------------------------------------------------------------------ our arm ----
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import ripoff_raytracer_b200 as rr
    from ripoff_raytracer_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the render path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()  # creates the NCCL communicator (and prints its banner) now
            torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    wl = make_workload(args)
    W, H, spp, bounces = wl.width, wl.height, wl.spp, wl.bounces
    tris, meshes, ranges, spheres = wl.scene.arrays()

    # host buffers of the boundary call live in pinned memory (torch owns the allocation)
    def pinned_copy(a):
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
        out = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)
        out[...] = a
        return t, out

    keep = []
    host = {}
    for name, arr in (("tris", tris), ("meshes", meshes), ("ranges", ranges), ("spheres", spheres)):
        t, v = pinned_copy(arr)
        keep.append(t)
        host[name] = v
    frame_t = torch.empty(W * H * 4, dtype=torch.uint8, pin_memory=True)
    frame_host = frame_t.numpy().reshape(H, W, 4)

    r = rr.Renderer((local,))
    r.upload_arrays(host["tris"], host["meshes"], host["ranges"], host["spheres"])

    # ---- multi-GPU plumbing: rank 0 owns the tile counter and the frame; peers attach over CUDA IPC ----
    mode = "local"
    if world > 1:
        from ripoff_raytracer_b200 import multigpu

        q, f = multigpu.exchange_handles(dist, rank, lambda: r.queue_export(W, H), device="cuda")
        attached = True
        if rank != 0:
            try:
                r.queue_import(W, H, q, f)
            except _abi.RRError as e:
                print(f"[rank {rank}] IPC import failed ({e}); falling back to a static tile partition", file=sys.stderr)
                attached = False
        mode = multigpu.negotiate_mode(dist, attached, device="cuda")

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def device_step():
        flush.fill_(1)  # evict L2 between steps
        if mode == "local":
            return r.render_device(wl.cam, W, H, spp, bounces)
        if mode == "shared":
            return shared_frame(r, wl.cam, W, H, spp, bounces, rank, barrier)
        st = r.render_strided(wl.cam, W, H, spp, bounces, rank, world)
        return st

    def total(x: float, op=None):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return float(t.item())

    # ---- per-ray test counts from one instrumented low-spp frame (same scene, same rays per pixel prefix) ----
    probe_spp = max(1, min(spp, 4))
    _, _, pst = r.render(wl.cam, W, H, probe_spp, bounces, count_tests=True)
    traced = max(pst["rays"], 1)
    n_box, n_tri, n_sph = pst["box_tests"] / traced, pst["tri_tests"] / traced, pst["sphere_tests"] / traced
    flops_per_ray = n_box * F_BOX + n_tri * F_TRI + n_sph * F_SPH
    bytes_per_ray = n_box * B_BOX + n_tri * B_TRI + n_sph * B_SPH

    # ---- warm-up, then K timed steps ----
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    rays_total = rays_traced = 0
    kernel_ms = []
    with ClockSampler(local) as clocks:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = device_step()
            rays_total += st["rays"]
            rays_traced += st["rays"]
            kernel_ms.append(st["render_ms"])  # CUDA events on the library's own stream, around the render kernel
        barrier()
        wall = time.perf_counter() - t0
    wall = total(wall, dist.ReduceOp.MAX if world > 1 else None)
    rays_all = total(float(rays_total))
    traced_all = total(float(rays_traced))
    # device time of a step = the slowest rank's kernel (all ranks start together behind the barrier)
    step_ms = torch.tensor(kernel_ms, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
    dt = float(step_ms.sum().item()) / 1e3
    kernel_s = dt / max(args.steps, 1)
    samples_all = float(W) * H * spp * args.steps
    value = rays_all / dt / 1e6
    clk = clocks.summary()

    # ---- end to end through the boundary call: host scene in, host frame out, every step ----
    e2e = None
    if not args.no_e2e:
        h2d = host["tris"].nbytes + host["meshes"].nbytes + host["spheres"].nbytes + host["ranges"].nbytes + 48
        d2h = W * H * 4
        e2e_steps = max(1, min(args.steps, 2))
        e_rays = 0.0
        e_kernel_ms = 0.0
        with ClockSampler(local) as e_clocks:
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                r.upload_arrays(host["tris"], host["meshes"], host["ranges"], host["spheres"])  # H2D + LBVH build
                if mode == "local":
                    _, _, st = r.render(wl.cam, W, H, spp, bounces, out=frame_host)              # render + D2H
                else:
                    st = device_step_e2e(r, rr, wl, mode, rank, world, barrier, frame_host)
                e_rays += st["rays"]
                e_kernel_ms += st["render_ms"]
            barrier()
            e_dt = total(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
        e_rays = total(e_rays)
        # "frame time to output.bmp" (BASELINE.json metric): the e2e step plus placeImageDataIntoBMP of the frame it left
        # in host memory (src/main.cpp:725); written once, outside the timed region, by rank 0
        bmp_ms = None
        if rank == 0:
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                t0 = time.perf_counter()
                rr.write_bmp(Path(td) / "output.bmp", frame_host)
                bmp_ms = (time.perf_counter() - t0) * 1e3
        e2e = {"value": e_rays / e_dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_dt / e2e_steps * 1e3, "steps": e2e_steps, "kernel_ms_per_step": e_kernel_ms / e2e_steps,
               "clocks": e_clocks.summary(),
               "includes": "rr_upload_scene (H2D from pinned memory + LBVH build) + rr_render (kernel + frame D2H)",
               "bmp_write_ms": bmp_ms,
               "frame_time_to_bmp_ms": (e_dt / e2e_steps * 1e3 + bmp_ms) if bmp_ms is not None else None}

    # ---- N > 1: is the frame the ranks render TOGETHER the frame one GPU renders alone?  (outside the timed regions) ----
    frame_equal = None
    if world > 1:
        chk_spp = 1
        if mode == "shared":
            shared_frame(r, wl.cam, W, H, chk_spp, bounces, rank, barrier)
            barrier()
            together = r.read_frame(W, H) if rank == 0 else None
        else:
            r.render_strided(wl.cam, W, H, chk_spp, bounces, rank, world)
            from ripoff_raytracer_b200 import multigpu as _mg
            together = _mg.merge_strided_frames(dist, r.read_frame(W, H), rank, device="cuda")
        if rank == 0:
            import zlib
            alone = rr.Renderer((local,))   # a context of its own: private queue, private frame
            alone.upload_arrays(host["tris"], host["meshes"], host["ranges"], host["spheres"])
            want, _, _ = alone.render(wl.cam, W, H, chk_spp, bounces)
            alone.close()
            frame_equal = {"equal": bool(np.array_equal(together, want)), "spp": chk_spp, "crc32_together": zlib.crc32(together.tobytes()),
                           "crc32_one_gpu": zlib.crc32(want.tobytes()), "differing_pixels": int((together != want).any(-1).sum())}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines of the dominant kernel (k_render): counted intersection work / live kernel time ----
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    import ctypes as C
    def probe_peak(what):
        v = C.c_float()
        _abi.check(_abi.lib().rr_probe_peak(what, C.byref(v)), "rr_probe_peak")
        return float(v.value)
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    sm_max = (clk.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
    fp32_nominal = sm_count * 128 * 2 * sm_max * 1e6 / 1e12  # 128 FP32 lanes/SM, FMA = 2 flop
    fp32_measured = probe_peak(0)                            # FFMA issue rate of this device, measured now
    l2_measured = probe_peak(1)                              # L2 read bandwidth of this device, measured now
    per_launch_rays = traced_all / args.steps / world
    ach_tflops = per_launch_rays * flops_per_ray / kernel_s / 1e12 if kernel_s else 0.0
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ach_gbs = per_launch_rays * bytes_per_ray / kernel_s / 1e9 if kernel_s else 0.0
    traffic, traffic_src = None, None
    try:  # dram__bytes_read+write of k_render from the committed `ncu --set full` capture, per traced ray
        ncu = json.loads((ROOT / "profiles" / "ncu_summary.json").read_text())
        per_ray = ncu["k_render"]["dram_bytes_per_ray"]
        traffic = per_ray * per_launch_rays
        traffic_src = f"{per_ray:.1f} B/ray x rays of this launch; measured by ncu on {ncu['k_render']['config']}"
    except Exception:
        pass
    # `roofline` is the roof north_star names and SURVEY.md 8d states for this path: counted intersection flops over the
    # FP32 FMA rate (measured on this device by rr_probe_peak; the nominal figure beside it).  The same launch against the
    # two memory roofs follows in `roofline_l2` and `roofline_bytes` (the counted bytes are served by L1 / L2: the DRAM
    # traffic ncu measured is a tenth of them).  None of the three binds a divergent BVH walk, which is bound by the
    # latency of its dependent node fetches: the useful efficiency figures are issue-slot
    # utilisation and lanes per instruction in profiles/.
    tests = {"box_tests_per_ray": n_box, "tri_tests_per_ray": n_tri, "sphere_tests_per_ray": n_sph}
    roofline = {
        "bound": "fp32", "kernel": "k_render", "achieved": ach_tflops, "peak": fp32_measured, "unit": "TFLOP/s",
        "frac": ach_tflops / fp32_measured if fp32_measured else None, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": "FP32 FMA rate measured on this device by rr_probe_peak(0) (MEASURED_PEAKS.json holds no FP32 figure)",
        "peak_nominal": fp32_nominal, "frac_of_nominal": ach_tflops / fp32_nominal,
        "flops_per_ray": flops_per_ray, **tests, "kernel_ms": kernel_s * 1e3,
        "note": "24 flop per box test, 53 per triangle test, 24 per sphere test (SURVEY.md 8d), counted by the instrumented kernel on "
                "the same scene; result arithmetic is issued unfused (numerics contract), slab tests use FFMA",
    }
    roofline_l2 = {"bound": "l2", "kernel": "k_render", "achieved": ach_gbs, "peak": l2_measured, "unit": "GB/s",
                   "frac": ach_gbs / l2_measured if l2_measured else None, "bytes_per_ray": bytes_per_ray,
                   "peak_source": "L2 read bandwidth measured on this device by rr_probe_peak(1): 32 MB read 64 times, 16-byte loads"}
    roofline_bytes = {"bound": "hbm", "kernel": "k_render", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                      "frac": ach_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                      "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
                      "bytes_per_ray": bytes_per_ray,
                      "note": "algorithmic bytes = 32 B per box test + 48 B per triangle test + 16 B per sphere test (SURVEY.md 8d); DRAM "
                              "moves far fewer (traffic): this is NOT the binding roof, the walk runs out of L1 / L2"}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(wl, world),
        "msamples_per_s": samples_all / dt / 1e6, "frame_time_s": dt / args.steps, "wall_ms_per_step": wall / args.steps * 1e3,
        "timing": "CUDA events on the launching stream around each render kernel, max over ranks per step",
        "rays_per_sample": rays_all / samples_all,
        "tile_queue": mode,
        "clocks": clk, "gpu_launches": args.steps * world, "roofline": roofline, "roofline_l2": roofline_l2,
        "roofline_bytes": roofline_bytes,
        # the instrumented probe frame walks the same hierarchy: traversal-stack overflows must be (and are) 0; a non-zero
        # count in any frame makes the library fail the render with RR_ERR_BVH_DEPTH
        "stack_overflows": int(pst["stack_overflows"]),
    }
    if frame_equal is not None:
        line["frame_equal"] = frame_equal["equal"]
        line["frame_check"] = frame_equal
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        try:
            res = cpu_reference_run(wl, 1, 1, budget_s=20.0)
            line["cpu_baseline"] = {"value": res["mrays_s"], "unit": "Mrays/s", "cores": res["cores"], "kind": res["kind"],
                                    "sample": res["sample"], "msamples_per_s": res["msamples_s"], "reference_sah": res["reference_sah"]}
            if res["reference_sah"]:  # next to our box tests per ray: the same count for the reference's own hierarchy
                line["roofline"]["box_tests_per_ray_reference_sah"] = res["reference_sah"]["box_tests_per_ray"]
                line["roofline"]["tri_tests_per_ray_reference_sah"] = res["reference_sah"]["tri_tests_per_ray"]
        except Exception as e:  # the checker is optional for the measurement itself
            line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def shared_frame(r, cam, W, H, spp, bounces, rank, barrier):
    """One frame through the shared tile queue, protocol of include/rr_api.h: every rank is back from the previous
    frame (barrier), rank 0 resets the counter, nobody pops before the reset is done (barrier), all render."""
    barrier()
    if rank == 0:
        r.queue_reset()
    barrier()
    return r.render_shared(cam, W, H, spp, bounces)


def device_step_e2e(r, rr, wl, mode, rank, world, barrier, frame_host):
    """One multi-GPU frame ending with the frame in rank 0's host buffer."""
    W, H = wl.width, wl.height
    if mode == "shared":
        st = shared_frame(r, wl.cam, W, H, wl.spp, wl.bounces, rank, barrier)
        barrier()
        if rank == 0:
            r.read_frame(W, H, out=frame_host)  # D2H straight into the pinned host frame
        return st
    import torch
    import torch.distributed as dist

    from ripoff_raytracer_b200 import multigpu

    st = r.render_strided(wl.cam, W, H, wl.spp, wl.bounces, rank, world)
    merged = multigpu.merge_strided_frames(dist, r.read_frame(W, H), rank, device="cuda")  # one NCCL reduce
    if rank == 0:
        frame_host[...] = merged
    return st


if __name__ == "__main__":
    main()
