#!/usr/bin/env python
"""BASELINE config 3: Cambridge 480x854 view rendered at the matcher's resolution (ds=8 -> 60x107 = 6420 points) with
render_novel_view, then NeRFMatch-Mini coarse matching of synthetic image features against the rendered 3D features.
The image backbone (timm ConvFormer) is out of scope, so image features are noisy copies of the point features."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfmatch_b200.config import build_renderer
from nerfmatch_b200.matching import coarse_match
from nerfmatch_b200.synthetic import K_CAMBRIDGE, make_pose, make_state_dict, make_unnorm_scene

dev = torch.device("cuda:0")
sd = make_state_dict(0, embed_type="mip", sigma_gain=120.0, sigma_bias=(8.0, 8.0))
ren = build_renderer(dict(embed="mip", nc=128, nf=128, app=False, white_bg=False, stop_layer=3, feat_comb="lin"), sd, precision="bf16", device=dev)
un = make_unnorm_scene(60.0)
c2w = torch.from_numpy(un @ make_pose(5))
K = torch.from_numpy(K_CAMBRIDGE)


def timed(fn, iters=10):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


ms_r, view = timed(lambda: ren.render_novel_view((480, 854), K, c2w, un, dev, downsample=8))
pt_feat = view["pt_feat"]
n = pt_feat.shape[0]
g = torch.Generator(dev).manual_seed(1)
perm = torch.randperm(n, device=dev, generator=g)
im_feat = pt_feat[perm] + 0.05 * pt_feat.std() * torch.randn(n, 256, device=dev, generator=g)
ms_m, out = timed(lambda: coarse_match(im_feat, pt_feat, 10.0, "mul", mutual=True, match_thres=0.0, return_conf=False))
i, j = out["match_ids"][1], out["match_ids"][2]
print(json.dumps(dict(workload="Cambridge 480x854, ds=8, mip 128+128, bf16 MLP + fp32 coarse matching", points=n,
                      render_ms=ms_r, rays_per_s=n / ms_r * 1e3, match_ms=ms_m, matches=int(out["pred_num"]),
                      correct_frac=float((perm[i] == j).float().mean()))))
