#!/usr/bin/env python
"""A/B harness for kernel experiments.

    python tools/ab.py build NAME[:-DFLAG[,-DFLAG...]] ...      (here, no GPU: nvcc cross-compiles variants/librr_NAME.so)
    python tools/ab.py run [--workloads c4,c1] [--spp 8] NAME[@ENV=VAL[,ENV=VAL]] ...   (on the GPU box)

`run` starts one process per variant (RR_B200_LIB selects the library) and prints one JSON line per variant and
workload: kernel-only Mrays/s (CUDA events inside the library, best of 3 after a warm-up frame).
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
VAR = ROOT / "variants"
CSRC = ROOT / "ripoff_raytracer_b200" / "csrc"


def build(specs):
    VAR.mkdir(exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition(":")
        extra = " ".join(flags.split(",")) if flags else ""
        out = VAR / f"librr_{name}.so"
        srcs = " ".join(str(CSRC / f) for f in ("rr_api.cu", "rr_lbvh.cu", "rr_render.cu", "rr_host.cpp"))
        cmd = (f"/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false "
               f"-Xcompiler -fPIC,-O2 {extra} -shared -o {out} {srcs}")
        print(cmd, flush=True)
        subprocess.run(cmd, shell=True, check=True)


def child(workloads_csv, spp):
    sys.path.insert(0, str(ROOT))
    import ripoff_raytracer_b200 as rr
    from ripoff_raytracer_b200 import workloads

    for name in workloads_csv.split(","):
        wl = workloads.WORKLOADS[name](width=1920, height=1080, spp=spp)
        r = rr.Renderer((0,))
        r.upload(wl.scene)
        r.render_device(wl.cam, wl.width, wl.height, 2, wl.bounces)
        best = min((r.render_device(wl.cam, wl.width, wl.height, wl.spp, wl.bounces) for _ in range(3)), key=lambda s: s["render_ms"])
        import zlib
        crc = zlib.crc32(r.read_frame(wl.width, wl.height).tobytes())  # same image <=> same checksum across variants
        print(json.dumps({"variant": os.environ.get("RR_AB_LABEL", "?"), "workload": wl.name, "ms": round(best["render_ms"], 3),
                          "mrays_s": round(best["rays"] / best["render_ms"] / 1e3, 1), "rays": best["rays"], "frame_crc32": crc}), flush=True)
        r.close()


def run(args):
    wls, spp = "c4", 8
    specs = []
    it = iter(args)
    for a in it:
        if a == "--workloads":
            wls = next(it)
        elif a == "--spp":
            spp = int(next(it))
        else:
            specs.append(a)
    for spec in specs:
        name, _, envs = spec.partition("@")
        env = dict(os.environ, RR_AB_LABEL=spec)
        if name != "default":
            env["RR_B200_LIB"] = str(VAR / f"librr_{name}.so")
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        subprocess.run([sys.executable, __file__, "child", wls, str(spp)], env=env, check=False)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "child":
        child(sys.argv[2], int(sys.argv[3]))
    else:
        run(sys.argv[2:])
