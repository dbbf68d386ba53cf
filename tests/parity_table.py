"""Parity table of the CUDA render path against the reference (TEST INFRASTRUCTURE: imports the oracle).

For every reference-generated golden (tests/golden/render_*.npz), a BASELINE config-1 view (120x160 = 19 200 rays of a
480x640 7-Scenes camera, oracle on the host cores) and a full-resolution 480x640 crop, in both precisions of the CUDA
path, the table holds per-ray statistics of every quantity the north star names:

  feat_*   cosine (min / 1st percentile / median) and rel-L2 of the composited features
  depth_*  per-ray relative error (median / 99th percentile / max) and rel-L2
  pts / pt3d   max and percentile |delta| in normalised units and in metres (un-normalised, scene scale 3 m)
  rgb      max |delta|;   sigma_coarse: rel-L2 / max |delta| of the raw density on the (identical) coarse samples
  idx      fraction of identical re-sampling indices

`python tests/parity_table.py --out gpurun_out/r02_parity.json` writes the table; tests/test_gpu_parity_table.py asserts
the bounds DESIGN.md states on it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from nmr_helpers import RENDER_CASES, build_renderer, case_setup, cosine_rows, load_golden, ray_rel_err, rel_l2  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402
from oracle.weights import K_7SCENES, make_pose, make_state_dict, make_unnorm_scene, uniform01  # noqa: E402

DEV = "cuda:0"
SCALE_M = 3.0      # make_unnorm_scene(): uniform scale 3 m -> "1 mm" = 3.3e-4 normalised units


def _pct(x, q):
    return float(np.percentile(x, q)) if len(x) else 0.0


def compare(out, ref, feat_lin=True):
    """out / ref: dicts of numpy arrays with the render keys (+ optional intermediates)."""
    row = {}
    for key in ("coarse", "fine"):
        if f"feat_{key}" in out and f"feat_{key}" in ref:
            cos = cosine_rows(out[f"feat_{key}"], ref[f"feat_{key}"])
            row[f"feat_{key}"] = dict(cos_min=float(cos.min()), cos_p1=_pct(cos, 1), cos_median=float(np.median(cos)),
                                      frac_cos_ge_0999=float((cos >= 0.999).mean()), rel_l2=rel_l2(out[f"feat_{key}"], ref[f"feat_{key}"]))
        d, dr = out[f"depth_{key}"].reshape(-1), ref[f"depth_{key}"].reshape(-1)
        e = np.abs(d - dr) / np.maximum(np.abs(dr), 1e-3)
        row[f"depth_{key}"] = dict(rel_median=float(np.median(e)), rel_p99=_pct(e, 99), rel_max=float(e.max()), rel_l2=rel_l2(d, dr))
        dp = np.linalg.norm(out[f"pts_{key}"] - ref[f"pts_{key}"], axis=-1)
        row[f"pts_{key}"] = dict(norm_median=float(np.median(dp)), norm_p99=_pct(dp, 99), norm_max=float(dp.max()),
                                 metres_median=float(np.median(dp)) * SCALE_M, metres_p99=_pct(dp, 99) * SCALE_M, metres_max=float(dp.max()) * SCALE_M,
                                 frac_within_1mm=float((dp * SCALE_M < 1e-3).mean()))
        row[f"rgb_{key}"] = dict(max_abs=float(np.abs(out[f"rgb_{key}"].reshape(-1, 3) - ref[f"rgb_{key}"].reshape(-1, 3)).max()))
    if "pt3d" in out and "pt3d" in ref:
        dp = np.linalg.norm(out["pt3d"] - ref["pt3d"], axis=-1)
        row["pt3d"] = dict(metres_median=float(np.median(dp)), metres_p99=_pct(dp, 99), metres_max=float(dp.max()),
                           frac_within_1mm=float((dp < 1e-3).mean()))
    if "raw_coarse" in out and "raw_coarse" in ref:
        s, sr = out["raw_coarse"][..., 3], ref["raw_coarse"][..., 3]
        row["sigma_coarse"] = dict(rel_l2=rel_l2(s, sr), max_abs=float(np.abs(s - sr).max()), ref_rms=float(np.sqrt(np.mean(sr.astype(np.float64) ** 2))))
        w, wr = out["weights_coarse"], ref["weights_coarse"]
        row["weights_coarse"] = dict(max_abs=float(np.abs(w - wr).max()))
    if "idx_fine" in out and "idx_fine" in ref:
        row["idx_fine_equal"] = float((out["idx_fine"] == ref["idx_fine"]).mean())
    if "z_fine" in out and "z_fine" in ref:
        dz = np.abs(out["z_fine"] - ref["z_fine"])
        row["z_fine"] = dict(max_abs=float(dz.max()), p99=_pct(dz.reshape(-1), 99), frac_le_2e5=float((dz <= 2e-5).mean()))
    return row


def render_cuda(ren, rays, t_rand=None, u_jitter=None, ray_id=None, unnorm=None):
    import torch

    T = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    r = T(rays)
    n = r.shape[0]
    eng = ren.engine(r.device)
    outs = eng.alloc_outputs(n, True, unnorm is not None, intermediates=True)
    if ray_id is None and ren.appearance_embedding:
        ray_id = np.ones(n, np.int64)
    args, keep = eng.make_args(n, r.shape[1], rays=r, ray_id=T(ray_id), t_rand=T(t_rand), u_jitter=T(u_jitter), flags=ren._flags(),
                               unnorm=unnorm, outs=outs)
    eng.render(args)
    torch.cuda.synchronize()
    return {k: v.detach().cpu().numpy() for k, v in outs.items()}


def emulated_oracle(cfg, sd, g, prec):
    """The reference algorithm with 16-bit tensor-core operands (oracle.mlp_forward_rounded): what ANY correct bf16 / fp16
    implementation of the MLP computes, up to fp32 accumulation order."""
    import functools

    rnd = O.round_fp16 if prec == "fp16" else O.round_bf16
    out = O.render_rays(cfg, sd, g["rays"], g.get("t_rand"), g.get("u_jitter"), g.get("ray_id"), keep_intermediates=True,
                        mlp=functools.partial(O.mlp_forward_rounded, rnd=rnd))
    out["pt3d"] = O.unnormalize_pts(out["pts_fine"], g["unnorm_scene"])
    return out


def golden_rows(precisions=("fp32", "bf16", "fp16")):
    """Per golden and precision: the CUDA path against the reference's fp32 output; for the 16-bit modes also the emulated
    16-bit-operand oracle against the reference (the error 16-bit operands cause by themselves) and the CUDA path against
    that emulated oracle (what the kernel adds on top: nothing but accumulation order)."""
    rows = {}
    for name in RENDER_CASES:
        g = load_golden(f"render_{name}.npz")
        cfg, sd, c = case_setup(g)
        for prec in precisions:
            ren = build_renderer(c, sd, precision=prec)
            out = render_cuda(ren, g["rays"], g.get("t_rand"), g.get("u_jitter"), g.get("ray_id"), g["unnorm_scene"])
            row = compare(out, g)
            row["n_rays"] = int(g["rays"].shape[0])
            row["sigma_gain"] = [float(x) for x in c["sigma_gain"]]
            row["feat_comb"] = c["feat_comb"]
            rows[f"golden/{name}/{prec}"] = row
            if prec != "fp32":
                emu = emulated_oracle(cfg, sd, g, prec)
                rows[f"golden/{name}/{prec}_emulated_oracle_vs_reference"] = dict(compare(emu, g), n_rays=row["n_rays"])
                rows[f"golden/{name}/{prec}_cuda_vs_emulated_oracle"] = dict(compare(out, emu), n_rays=row["n_rays"])
    return rows


def view_case(embed, ds, crop=None, seed=31, gain=120.0, bias=(8.0, 8.0), nc=64, nf=128):
    """A seeded 7-Scenes view: rays, draws, oracle render (keep_intermediates) and the case dict."""
    rays = O.sample_nerf_rays(480, 640, K_7SCENES, make_pose(seed), ds=ds, embed_type=embed)
    if crop is not None:
        h, w = len(range(ds // 2, 480, ds)), len(range(ds // 2, 640, ds))
        y0, x0, ch, cw = crop
        rays = np.ascontiguousarray(rays.reshape(h, w, -1)[y0:y0 + ch, x0:x0 + cw].reshape(ch * cw, -1))
    n = rays.shape[0]
    tr = uj = None
    if embed == "mip":
        tr = uniform01(seed, 1, n * (nc + 1)).reshape(n, nc + 1)
        uj = (uniform01(seed, 2, n * (nc + 1)).reshape(n, nc + 1) * (1 / (nc + 1) - 1.2e-7)).astype(np.float32)
    sd = make_state_dict(seed, embed_type=embed, sigma_gain=gain, sigma_bias=bias)
    cfg = O.RenderConfig(embed_type=embed, num_pts_coarse=nc, num_pts_fine=nf, stop_layer=3)
    c = dict(embed=embed, nc=nc, nf=nf, app=False, stop_layer=3, feat_comb="lin", white_bg=False)
    return rays, tr, uj, sd, cfg, c


def oracle_render(cfg, sd, rays, tr, uj, unnorm):
    """Oracle on all host cores (numpy releases the GIL in BLAS; chunks of 256 rays per task)."""
    from concurrent.futures import ThreadPoolExecutor

    ch = 256
    n = rays.shape[0]

    def work(i):
        return O.render_rays(cfg, sd, rays[i:i + ch], None if tr is None else tr[i:i + ch], None if uj is None else uj[i:i + ch],
                             keep_intermediates=True, chunk=ch)

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        parts = list(ex.map(work, range(0, n, ch)))
    ref = {k: np.concatenate([p[k] for p in parts], 0) for k in parts[0]}
    ref["pt3d"] = O.unnormalize_pts(ref["pts_fine"], unnorm)
    return ref


def view_rows(precisions=("fp32", "bf16", "fp16"), quick=False, specs=None):
    rows = {}
    un = make_unnorm_scene()
    specs = specs or {
        # BASELINE config 1: one 7-Scenes view at 120x160 (480x640 camera, every 4th pixel), 64 + 128 samples
        "config1_120x160/normal": dict(embed="normal", ds=4),
        "config1_120x160/mip": dict(embed="mip", ds=4),
        # BASELINE config 2 resolution: a 64 x 96 crop of the full-resolution 480x640 ray grid (adjacent pixels)
        "config2_crop_64x96/normal": dict(embed="normal", ds=1, crop=(208, 272, 64, 96)),
        "config2_crop_64x96/mip128": dict(embed="mip", ds=1, crop=(208, 272, 64, 96), nc=128, nf=128),
    }
    if quick:
        specs = {"config1_120x160/normal": dict(embed="normal", ds=16), "config1_120x160/mip": dict(embed="mip", ds=16)}
    for name, sp in specs.items():
        rays, tr, uj, sd, cfg, c = view_case(**sp)
        ref = oracle_render(cfg, sd, rays, tr, uj, un)
        for prec in precisions:
            ren = build_renderer(c, sd, precision=prec)
            out = render_cuda(ren, rays, tr, uj, None, un)
            row = compare(out, ref)
            row["n_rays"] = int(rays.shape[0])
            row["sigma_gain"] = [120.0, 120.0]
            rows[f"{name}/{prec}"] = row
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_parity.json"))
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    rows = golden_rows()
    rows.update(view_rows(quick=a.quick))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)
    for k, r in rows.items():
        ff = r.get("feat_fine", {})
        print(f"{k:46s} n={r['n_rays']:6d} cos_f min {ff.get('cos_min', float('nan')):.5f} p1 {ff.get('cos_p1', float('nan')):.5f} med {ff.get('cos_median', float('nan')):.6f} "
              f"| depth_f med {r['depth_fine']['rel_median']:.2e} max {r['depth_fine']['rel_max']:.2e} | pt3d max {r.get('pt3d', {}).get('metres_max', float('nan')):.2e} m "
              f"p99 {r.get('pt3d', {}).get('metres_p99', float('nan')):.2e} | sig_c rel {r.get('sigma_coarse', {}).get('rel_l2', float('nan')):.2e} | idx {r.get('idx_fine_equal', float('nan')):.4f}")


if __name__ == "__main__":
    main()
