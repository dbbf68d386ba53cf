"""Configuration helpers: the nested ``argparse.Namespace`` the reference builds from
configs/nerf/*.yaml (utils/config.py:26-33), with the shipped render-relevant defaults
(configs/nerf/nerf_7scenes_mip_sfm.yaml:24-55)."""
from __future__ import annotations

from argparse import Namespace as NS


def make_config(embed="mip", nc=64, nf=128, app=False, white_bg=False, xyz_num_freqs=15, dirs_num_freqs=4, use_disp=False,
                mip_var_scale=-1):
    emb = NS(xyz_num_freqs=xyz_num_freqs, dirs_num_freqs=dirs_num_freqs, appearance_embed=app, type=embed, mip_var_scale=mip_var_scale)
    return NS(
        data=NS(img_wh=[160, 120]),
        render=NS(chunksize=16384, use_disp=use_disp, perturb=True, noise_std=1.0, white_bg=white_bg, use_viewdirs=True),
        embedding=emb,
        coarse_nerf=NS(method="NeRF", layer_num=8, hid_dim=256, output_dim=4, skips=[4], num_pts=nc),
        fine_nerf=NS(method="NeRF", layer_num=8, hid_dim=256, output_dim=4, skips=[4], num_pts=nf),
        loss=NS(),
    )


def build_renderer(case, state_dict, precision="bf16", device="cuda:0", num_frames=7, **kw):
    """NerfRenderer for a case dict {embed, nc, nf, app, white_bg, stop_layer, feat_comb} with a
    (numpy or torch) reference-keyed state_dict loaded - the construction sequence of
    nerf_evaluator.py:119-146."""
    import numpy as np
    import torch

    from .renderer import NerfRenderer

    ren = NerfRenderer(make_config(case["embed"], case["nc"], case["nf"], case["app"], case["white_bg"],
                                   use_disp=case.get("use_disp", False), mip_var_scale=case.get("mip_var_scale", -1)),
                       num_frames=num_frames if case["app"] else None, training=False, stop_layer=case["stop_layer"],
                       precision=precision, **kw)
    ren.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}, strict=True)
    ren = ren.to(device).eval()
    ren.ret_pfeat = True
    ren.feat_comb = case["feat_comb"]
    return ren
