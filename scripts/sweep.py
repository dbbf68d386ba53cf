#!/usr/bin/env python
"""Throughput sweep of the feature render (BASELINE config 5): rays in {1e5 .. 1e8} x samples in {64, 128, 256}
(classic sampling: Sc coarse + Sc fine extra samples, i.e. 3*Sc MLP evaluations per ray ... see table), plus
the Cambridge 480x854 view (config 3's render half).  Writes one JSON object per line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfmatch_b200 import _lib as L
from nerfmatch_b200.config import build_renderer
from nerfmatch_b200.synthetic import K_7SCENES, K_CAMBRIDGE, make_pose, make_state_dict


def time_render(eng, args, iters):
    for _ in range(2):
        eng.render(args)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        eng.render(args)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-rays", type=float, default=1e7)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    cases = []
    for embed, nc, nf in (("normal", 64, 128), ("mip", 64, 64), ("mip", 128, 128), ("normal", 128, 128)):
        for n in (1e5, 1e6, 1e7, 1e8):
            if n <= a.max_rays:
                cases.append((embed, nc, nf, int(n), "7scenes"))
    cases.append(("normal", 64, 128, 480 * 854, "cambridge_480x854"))
    cases.append(("mip", 128, 128, 480 * 854, "cambridge_480x854"))
    rens = {}
    for embed, nc, nf, n, tag in cases:
        key = (embed, nc, nf)
        if key not in rens:
            sd = make_state_dict(0, embed_type=embed, sigma_gain=120.0, sigma_bias=(8.0, 8.0))
            c = dict(embed=embed, nc=nc, nf=nf, app=False, stop_layer=3, feat_comb="lin", white_bg=False)
            rens[key] = build_renderer(c, sd, precision="bf16", device=dev)
        ren = rens[key]
        eng = ren.engine(dev)
        if tag.startswith("cambridge"):
            rays, _, _ = eng.raygen(K_CAMBRIDGE, make_pose(5), 480, 854, 1)
        else:
            base, _, _ = eng.raygen(K_7SCENES, make_pose(5), 480, 640, 1)
            reps = -(-n // base.shape[0])
            rays = base.repeat(reps, 1)[:n].contiguous() if reps > 1 else base[:n].contiguous()
        n = rays.shape[0]
        # large sweeps: only the matcher-facing outputs (features + points), like a streaming consumer would keep
        outs = {"feat_fine": torch.empty(n, 256, device=dev), "pts_fine": torch.empty(n, 3, device=dev)}
        args, keep = eng.make_args(n, rays.shape[1], rays=rays, flags=L.NMR_FLAG_RET_PFEAT | L.NMR_FLAG_SKIP_COARSE_FEAT,
                                   rng_mode=L.NMR_RNG_PHILOX if embed == "mip" else L.NMR_RNG_NONE, seed=1, outs=outs)
        iters = 3 if n <= 1e6 else 1
        ms = time_render(eng, args, iters)
        evals = (nc + nc) if embed == "mip" else (nc + nc + nf)
        row = dict(workload=tag, embed=embed, coarse=nc, fine=nf, mlp_evals_per_ray=evals, rays=n, ms=ms, rays_per_s=n / ms * 1e3,
                   tflops=n * evals * (1214464 if embed == "mip" else 1217536) / ms / 1e9)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del outs, rays
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
