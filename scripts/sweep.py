#!/usr/bin/env python
"""Throughput sweep of the feature render (BASELINE config 5): rays in {1e5 .. 1e8} x samples per ray, at 1/2/4/8 GPUs
(under torchrun every rank renders its contiguous share of the rays; time = max over ranks), plus the Cambridge 480x854 view
(config 3's render half).  Above 1e7 rays per GPU the outputs are streamed: the rays are rendered in batches of 1e7 into the
same output buffers, like a consumer that hands each batch on (the features alone would be 100 GB at 1e8 rays).
Writes one JSON object per line; `cpu_rays_per_s` (the numpy port of the reference on the host cores, measured on a bounded
sample by bench.py and linear in the ray count) can be passed in to get the ratio column."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfmatch_b200 import _lib as L
from nerfmatch_b200.config import build_renderer
from nerfmatch_b200.dist import shard_range
from nerfmatch_b200.synthetic import K_7SCENES, K_CAMBRIDGE, make_pose, make_state_dict

BATCH = 10_000_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-rays", type=float, default=1e8)
    ap.add_argument("--min-rays", type=float, default=1e5)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--cpu-rays-per-s", type=float, default=0.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cases = []
    for embed, nc, nf in (("mip", 64, 64), ("normal", 64, 128), ("mip", 128, 128), ("normal", 128, 128)):
        for n in (1e5, 1e6, 1e7, 1e8):
            if a.min_rays <= n <= a.max_rays:
                cases.append((embed, nc, nf, int(n), "7scenes"))
    if world == 1:
        cases += [("normal", 64, 128, 480 * 854, "cambridge_480x854"), ("mip", 128, 128, 480 * 854, "cambridge_480x854")]
    rens, rows = {}, []
    for embed, nc, nf, n_total, tag in cases:
        key = (embed, nc, nf)
        if key not in rens:
            sd = make_state_dict(0, embed_type=embed, sigma_gain=120.0, sigma_bias=(8.0, 8.0))
            c = dict(embed=embed, nc=nc, nf=nf, app=False, stop_layer=3, feat_comb="lin", white_bg=False)
            rens[key] = build_renderer(c, sd, precision=a.precision, device=dev)
        eng = rens[key].engine(dev)
        s, e = shard_range(n_total, world, rank)
        n = e - s                                                   # this rank's rays
        nb = min(n, BATCH)
        if tag.startswith("cambridge"):
            base, _, _ = eng.raygen(K_CAMBRIDGE, make_pose(5), 480, 854, 1)
        else:
            base, _, _ = eng.raygen(K_7SCENES, make_pose(5 + rank), 480, 640, 1)
        reps = -(-nb // base.shape[0])
        rays = base.repeat(reps, 1)[:nb].contiguous() if reps > 1 else base[:nb].contiguous()
        # large sweeps: only the matcher-facing outputs (features + points), like a streaming consumer would keep
        outs = {"feat_fine": torch.empty(nb, 256, device=dev), "pts_fine": torch.empty(nb, 3, device=dev)}
        flags = L.NMR_FLAG_RET_PFEAT | L.NMR_FLAG_SKIP_COARSE_FEAT
        rng = L.NMR_RNG_PHILOX if embed == "mip" else L.NMR_RNG_NONE
        full, keep = eng.make_args(nb, rays.shape[1], rays=rays, flags=flags, rng_mode=rng, seed=1, outs=outs)
        batches = [full] * (n // nb)
        if n % nb:
            tail, keep2 = eng.make_args(n % nb, rays.shape[1], rays=rays, flags=flags, rng_mode=rng, seed=1, outs=outs)
            batches.append(tail)

        def run():
            for b in batches:
                eng.render(b)

        eng.render(full if nb <= 2_000_000 else batches[-1])        # warm-up (workspace allocation)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        iters = 3 if n_total <= 1e6 else 1
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(iters):
            run()
        ev1.record()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / iters], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        evals = (nc + nc) if embed == "mip" else (nc + nc + nf)
        row = dict(workload=tag, n_gpus=world, precision=a.precision, embed=embed, coarse=nc, fine=nf, mlp_evals_per_ray=evals, rays=n_total, ms=ms,
                   rays_per_s=n_total / ms * 1e3, tflops=n_total * evals * (1214464 if embed == "mip" else 1217536) / ms / 1e9,
                   streamed_batches=len(batches))
        if a.cpu_rays_per_s > 0:
            row["cpu_reference_s_extrapolated"] = n_total / a.cpu_rays_per_s * (evals / 256.0)
            row["speedup_vs_cpu_port"] = row["cpu_reference_s_extrapolated"] / (ms / 1e3)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del outs, rays, full, batches
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
