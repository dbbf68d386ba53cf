#!/usr/bin/env python
"""BASELINE config 4: a batch of 64 query poses rendered at matcher resolution (480x640, ds=8 -> 4800 points each),
rays sharded across the ranks, pt_feat / pt3d / im_pred gathered to rank 0.  Launch with torchrun (or plain python at N=1).
Prints one JSON line on rank 0: views/s and rays/s, device timed (CUDA events, max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfmatch_b200.config import build_renderer
from nerfmatch_b200.dist import render_views_sharded
from nerfmatch_b200.synthetic import K_7SCENES, make_pose, make_state_dict, make_unnorm_scene

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
else:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
sd = make_state_dict(0, embed_type="normal", sigma_gain=120.0, sigma_bias=(8.0, 8.0))
ren = build_renderer(dict(embed="normal", nc=64, nf=128, app=False, white_bg=False, stop_layer=3, feat_comb="lin"), sd, precision="bf16", device=dev)
un = make_unnorm_scene(2.0)
V = 64
c2ws = torch.from_numpy(np.stack([un @ make_pose(i) for i in range(V)]))
K = torch.from_numpy(K_7SCENES)


def step():
    return render_views_sharded(ren, (480, 640), K, c2ws, un, downsample=8, dst=0)


for _ in range(3):
    out = step()
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
a.record()
for _ in range(iters):
    out = step()
b.record()
torch.cuda.synchronize(); dist.barrier()
ms = torch.tensor([a.elapsed_time(b) / iters], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    n = V * out["pt_feat"].shape[1]
    print(json.dumps(dict(workload="64 query poses, 480x640 ds=8 (4800 points each), classic 64+128, bf16, gather to rank 0", n_gpus=world,
                          ms_per_batch=ms.item(), views_per_s=V / ms.item() * 1e3, rays_per_s=n / ms.item() * 1e3,
                          finite=bool(torch.isfinite(out["pt_feat"]).all()), shape=list(out["pt_feat"].shape))))
dist.destroy_process_group()
