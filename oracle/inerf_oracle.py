"""TEST INFRASTRUCTURE ONLY - CPU oracle for the differentiable render of the iNeRF pose refinement
(not built as a CUDA path yet - this file and tests/golden/inerf_grad.npz pin what it has to compute).

Restates, in numpy float64, the graph the reference differentiates in one refinement step
(nerfmatch/nerfmatch_evaluator.py):
  :236-285  gen_rays: rays from the pose being optimised (origin = t, viewdirs = normalize(R @ K^-1 [x, y, 1]))
  :345-376  sample positions are taken from DETACHED rays (z and the Gaussian variances are constants);
            the points are rebuilt as  o + t_mean(z) * viewdirs  so that they depend on the pose
  :377-404  IPE of (points, constant variances), PE of the view directions, fine NeRF MLP
  :410-424  volume_render_radiance_field(white_bg=True, mip intervals) and the MSE photometric loss
and its analytic backward down to d loss / d pose.  Pinned against torch autograd through the reference's own
functions (oracle/gen_golden.py::inerf_cases)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

F64 = np.float64
HALF_PI32 = F64(np.float32(0.5 * np.pi))      # the reference adds fp32(0.5*pi) (embedding.py:69)
EPS32 = F64(np.finfo(np.float32).eps)


def _pixel_dirs(H: int, W: int, K: np.ndarray, ds: int) -> np.ndarray:
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    xys = np.stack([xs, ys, np.ones_like(xs)], -1).astype(F64)
    dirs = xys @ np.linalg.inv(K.astype(F64)).T
    return dirs[ds // 2 :: ds, ds // 2 :: ds].reshape(-1, 3)


def _t_mean(z: np.ndarray) -> np.ndarray:
    mu = (z[:, :-1] + z[:, 1:]) / 2
    hw = (z[:, :-1] - z[:, 1:]) / 2
    return mu + (2 * mu * hw**2) / np.maximum(EPS32, 3 * mu**2 + hw**2)


def fine_pass_loss_and_grad(pose: np.ndarray, K: np.ndarray, H: int, W: int, ds: int, z_vals: np.ndarray, var: np.ndarray,
                            wts: Dict[str, np.ndarray], img: np.ndarray, xyz_freqs: int = 15, dirs_freqs: int = 4,
                            skips=(4,)) -> Tuple[float, np.ndarray, Dict[str, np.ndarray]]:
    """loss, d loss / d pose [4,4] (rows 0..2 only), intermediates {rgb_map, grad_raw [n*S,4], grad_inputs [n*S,117]}.
    ``z_vals`` [n,S+1] and ``var`` [n,S,3] are the (constant) fine-pass samples; ``wts`` the fine net's state_dict."""
    pose = pose.astype(F64)
    R, t = pose[:3, :3], pose[:3, 3]
    dirs = _pixel_dirs(H, W, K, ds)                               # [n,3]
    n, S = z_vals.shape[0], z_vals.shape[1] - 1
    z = z_vals.astype(F64)
    # ---- forward -------------------------------------------------------------------------------------------------
    d = dirs @ R.T
    dn = np.sqrt(np.sum(d * d, -1, keepdims=True))
    v = d / dn                                                    # viewdirs = rays_d
    tm = _t_mean(z)                                               # [n,S]
    pts = t[None, None, :] + tm[..., None] * v[:, None, :]        # [n,S,3]
    sx = 2.0 ** np.arange(xyz_freqs)
    arg = (pts[:, :, None, :] * sx[None, None, :, None]).reshape(n, S, -1)               # [n,S,45] freq-major, xyz-minor
    damp = np.exp(-0.5 * (var.astype(F64)[:, :, None, :] * (sx * sx)[None, None, :, None]).reshape(n, S, -1))
    enc = np.concatenate([damp * np.sin(arg), damp * np.sin(arg + HALF_PI32)], -1)       # [n,S,90]
    sd_ = 2.0 ** np.arange(dirs_freqs)
    varg = (v[:, None, :] * sd_[None, :, None]).reshape(n, -1)                           # [n,12]
    de = np.concatenate([np.sin(varg), np.sin(varg + HALF_PI32), v], -1)                 # [n,27]
    x = np.concatenate([enc, np.broadcast_to(de[:, None, :], (n, S, de.shape[-1]))], -1).reshape(n * S, -1)
    xd = enc.shape[-1]
    W_ = lambda k: wts[k].astype(F64)
    in_pts, in_views = x[:, :xd], x[:, xd:]
    acts, h = [], in_pts
    for i in range(8):
        pre = h @ W_(f"pts_linears.{i}.weight").T + W_(f"pts_linears.{i}.bias")
        acts.append((h, pre))
        h = np.maximum(pre, 0)
        if i in skips:
            h = np.concatenate([in_pts, h], -1)
    h8 = h
    sigma_raw = h8 @ W_("alpha_linear.weight").T + W_("alpha_linear.bias")              # [P,1]
    feature = h8 @ W_("feature_linear.weight").T + W_("feature_linear.bias")
    hv_in = np.concatenate([feature, in_views], -1)
    hv_pre = hv_in @ W_("views_linears.0.weight").T + W_("views_linears.0.bias")
    hv = np.maximum(hv_pre, 0)
    rgb_lin = hv @ W_("rgb_linear.weight").T + W_("rgb_linear.bias")
    rgb = 1 / (1 + np.exp(-rgb_lin))                                                      # [P,3]
    # volume_render_radiance_field (render_utils.py:176-230), mip intervals, rays_d = viewdirs
    c = rgb.reshape(n, S, 3)
    sg = np.maximum(sigma_raw.reshape(n, S), 0)
    vn = np.sqrt(np.sum(v * v, -1, keepdims=True))                                        # = 1, but differentiated
    delta = (z[:, 1:] - z[:, :-1]) * vn
    ex = np.exp(-sg * delta)
    alpha = 1 - ex
    f = 1 - alpha + 1e-10
    T = np.concatenate([np.ones((n, 1)), np.cumprod(f, -1)[:, :-1]], -1)
    w = alpha * T
    acc = w.sum(-1)
    rgb_map = (w[..., None] * c).sum(1) + (1 - acc)[:, None]                              # white background
    diff = rgb_map - img.astype(F64)
    loss = float(np.mean(diff * diff))
    # ---- backward ------------------------------------------------------------------------------------------------
    g_map = 2 * diff / diff.size                                                          # [n,3]
    g_c = w[..., None] * g_map[:, None, :]
    g_w = (g_map[:, None, :] * (c - 1)).sum(-1)                                           # incl. the white background
    suffix = np.cumsum((g_w * w)[:, ::-1], -1)[:, ::-1] - g_w * w                         # sum_{k>s} g_w_k w_k
    g_alpha = g_w * T - suffix / f
    g_sg = g_alpha * ex * delta
    g_delta = g_alpha * ex * sg
    g_vn = (g_delta * (z[:, 1:] - z[:, :-1])).sum(-1, keepdims=True)
    g_v = g_vn * v / vn
    g_sigma_raw = (g_sg * (sigma_raw.reshape(n, S) > 0)).reshape(-1, 1)
    g_rgb_lin = (g_c.reshape(-1, 3)) * rgb * (1 - rgb)
    grad_raw = np.concatenate([g_c.reshape(-1, 3), g_sigma_raw], -1)      # w.r.t. the MLP output [sigmoid rgb, raw sigma]
    g_hv = g_rgb_lin @ W_("rgb_linear.weight")
    g_hv_pre = g_hv * (hv_pre > 0)
    g_hv_in = g_hv_pre @ W_("views_linears.0.weight")
    g_feature, g_views = g_hv_in[:, :256], g_hv_in[:, 256:]
    g_h = g_feature @ W_("feature_linear.weight") + g_sigma_raw @ W_("alpha_linear.weight")
    g_in_pts = np.zeros_like(in_pts)
    for i in reversed(range(8)):
        if i in skips:                                           # h = cat(in_pts, relu(pre))
            g_in_pts += g_h[:, :xd]
            g_h = g_h[:, xd:]
        g_pre = g_h * (acts[i][1] > 0)
        g_h = g_pre @ W_(f"pts_linears.{i}.weight")
    g_in_pts += g_h
    grad_inputs = np.concatenate([g_in_pts, g_views], -1)
    # encodings
    g_enc = g_in_pts.reshape(n, S, -1)
    nb = arg.shape[-1]
    g_arg = g_enc[..., :nb] * damp * np.cos(arg) + g_enc[..., nb:] * damp * np.cos(arg + HALF_PI32)
    g_pts = (g_arg.reshape(n, S, xyz_freqs, 3) * sx[None, None, :, None]).sum(2)          # [n,S,3]
    g_de = g_views.reshape(n, S, -1).sum(1)                                               # dirs encoding is shared by the ray's samples
    nd = varg.shape[-1]
    g_varg = g_de[:, :nd] * np.cos(varg) + g_de[:, nd : 2 * nd] * np.cos(varg + HALF_PI32)
    g_v = g_v + (g_varg.reshape(n, dirs_freqs, 3) * sd_[None, :, None]).sum(1) + g_de[:, 2 * nd :]
    # points = t + t_mean * v
    g_t = g_pts.sum((0, 1))
    g_v = g_v + (g_pts * tm[..., None]).sum(1)
    # v = d / |d| ; d = dirs @ R^T
    g_d = (g_v - v * np.sum(v * g_v, -1, keepdims=True)) / dn
    g_R = g_d.T @ dirs
    grad_pose = np.zeros((4, 4), F64)
    grad_pose[:3, :3] = g_R
    grad_pose[:3, 3] = g_t
    return loss, grad_pose, {"rgb_map": rgb_map, "grad_raw": grad_raw, "grad_inputs": grad_inputs}
