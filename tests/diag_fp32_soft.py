"""Diagnostic (not a test): which rays of the fp32-path-vs-oracle comparisons exceed 1e-3, and why."""
import sys, os
import numpy as np
import torch
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path[:0] = [os.path.dirname(HERE), HERE]
from nmr_helpers import build_renderer, ray_rel_err
from oracle import nerf_oracle as O
from oracle.weights import K_7SCENES, make_pose, make_state_dict, uniform01
import parity_table as PT

def run(embed, seed_w, pose, ds, extra, okw, seed_d):
    c = dict(embed=embed, nc=64, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False, **extra)
    sd = make_state_dict(seed_w, embed_type=embed, sigma_gain=120.0, sigma_bias=(8.0, 8.0))
    rays = O.sample_nerf_rays(480, 640, K_7SCENES, make_pose(pose), ds=ds, embed_type=embed)
    n = rays.shape[0]
    tr = uj = None
    if embed == "mip":
        tr = uniform01(seed_d, 1, n * 65).reshape(n, 65)
        uj = (uniform01(seed_d, 2, n * 65).reshape(n, 65) * (1 / 65 - 1.2e-7)).astype(np.float32)
    cfg = O.RenderConfig(embed_type=embed, num_pts_coarse=64, num_pts_fine=128, stop_layer=3, **okw)
    ref = O.render_rays(cfg, sd, rays, tr, uj, keep_intermediates=True)
    out = PT.render_cuda(build_renderer(c, sd, precision="fp32"), rays, tr, uj)
    if embed == "normal":
        zc, wc = ref["z_coarse"], ref["weights_coarse"]
        w = (wc[:, 1:-1] + np.float32(1e-5)); pdf = w / w.sum(-1, keepdims=True); cdf = np.concatenate([np.zeros((n, 1), np.float32), np.cumsum(pdf, -1)], -1)
        inds = ref["idx_fine"].astype(np.int64)
        below, above = np.maximum(0, inds - 1), np.minimum(cdf.shape[-1] - 1, inds)
        denom = np.take_along_axis(cdf, above, 1) - np.take_along_axis(cdf, below, 1)
        mind = denom.min(-1)
        e_all = np.maximum.reduce([ray_rel_err(out[f"{q}_fine"], ref[f"{q}_fine"]) for q in ("feat", "pts", "rgb", "depth")])
        for tau in (1e-5, 1e-4, 1e-3, 1e-2):
            good = mind >= tau
            print("  tau", tau, "well-conditioned rays", int(good.sum()), "max e among them", float(e_all[good].max()) if good.any() else None,
                  "max e among the rest", float(e_all[~good].max()) if (~good).any() else None)
    for q in ("feat", "pts", "rgb", "depth"):
        e = ray_rel_err(out[f"{q}_fine"], ref[f"{q}_fine"])
        bad = e >= 1e-3
        s_ref, s_out = ref["raw_fine"][:, -1, 3], out["raw_fine"][:, -1, 3]
        flip = (s_ref > 0) != (s_out > 0)
        dz = np.abs(out["z_fine"] - ref["z_fine"]).max(-1)
        zbad = dz > 1e-6
        wl_ref = ref["weights_fine"][:, -1]; wl_out = out["weights_fine"][:, -1]
        print(embed, extra, q, "n", n, "bad", int(bad.sum()), "of which last-sigma sign flips", int((bad & flip).sum()), "z_fine differs", int((bad & zbad).sum()),
              "neither", int((bad & ~flip & ~zbad).sum()), "| all flips", int(flip.sum()), "all zdiff", int(zbad.sum()), "median e", float(np.median(e)), "max e good", float(e[~flip & ~zbad].max()) if (~flip & ~zbad).any() else None)
        if q == "depth":
            idx = np.where(bad)[0][:6]
            for i in idx:
                print("   ray", i, "e", e[i], "sig_last ref/out", s_ref[i], s_out[i], "w_last ref/out", wl_ref[i], wl_out[i], "dz", dz[i], "idx_eq", float((out["idx_fine"][i] == ref["idx_fine"][i]).mean()))

run("mip", 41, 7, 16, {}, {}, 42)
run("normal", 41, 7, 16, {}, {}, 42)
run("normal", 43, 9, 32, dict(use_disp=True), dict(use_disp=True), 44)
run("mip", 43, 9, 32, dict(mip_var_scale=2.5), dict(mip_var_scale=2.5), 44)
