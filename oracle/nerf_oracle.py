"""CPU oracle for the NeRFMatch scene-feature render path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy (fp32) restatement of the reference algorithm in
``nerfmatch/nerf`` (ray generation -> stratified / hierarchical sampling ->
positional encoding -> NeRF MLP -> alpha compositing of features / depth / 3D points).
Nothing under ``nerfmatch_b200/`` may import it: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do, as the checker.

Parity status: the reference ships NO tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
generated in the build container by ``oracle/gen_golden.py`` (imports /root/reference)
and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every
function below against those fixtures.

Conventions
  * all arithmetic is float32 unless stated; the two scans (cumsum / cumprod) accumulate
    in float64 and round every prefix to float32, which is what torch's CPU kernels do
    (probed: torch.cumsum/cumprod on CPU == float64-accumulate bit for bit);
  * row sums that feed the CDF (``weight_sum``) are accumulated sequentially in float64
    and rounded once - torch's vectorised CPU sum order is build dependent, so this is
    the order the CUDA path reproduces bit for bit (DESIGN.md "bit-exact indices");
  * random draws are never made here: ``t_rand`` / ``u_jitter`` are inputs.

Every function cites the reference file:line it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

F32 = np.float32
EPS32 = float(np.finfo(np.float32).eps)


# ----------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------
@dataclass
class RenderConfig:
    """Subset of the reference YAML that the render path reads (renderer.py:27-114)."""

    embed_type: str = "mip"  # "mip" | "normal"   (renderer.py:40)
    xyz_num_freqs: int = 15
    dirs_num_freqs: int = 4
    num_pts_coarse: int = 64
    num_pts_fine: int = 128
    hid_dim: int = 256
    layer_num: int = 8
    skips: Tuple[int, ...] = (4,)
    use_viewdirs: bool = True
    app_dim: int = 0  # 16 when embedding.appearance_embed (renderer.py:72-73)
    stop_layer: int = 3  # FINE net feature tap (models/nerf.py:103-104); <0 -> last trunk layer
    stop_layer_coarse: int = -1  # renderer.py:99 sets stop_layer on fine_conf only -> coarse taps h7
    feat_comb: str = "lin"  # "lin" | "max" (renderer.py:257-265)
    white_bg: bool = False
    use_disp: bool = False
    mip_var_scale: float = -1.0
    resample_padding: float = 0.01
    ret_pfeat: bool = True

    @property
    def xyz_dim(self) -> int:
        # mip drops the raw xyz (renderer.py:56); Fourier keeps it (embedding.py:32-33)
        return 6 * self.xyz_num_freqs + (0 if self.embed_type == "mip" else 3)

    @property
    def dirs_dim(self) -> int:
        return (6 * self.dirs_num_freqs + 3) if self.use_viewdirs else 0


# ----------------------------------------------------------------------------------
# small helpers with torch-CPU semantics
# ----------------------------------------------------------------------------------
def torch_linspace01(steps: int) -> np.ndarray:
    """torch.linspace(0, 1, steps) in fp32 on CPU: lower half ``i*step``, upper half
    ``fma(-step, steps-1-i, 1)`` (probed against torch 2.11; used by
    render_utils.py:110, 244, 434)."""
    if steps == 1:
        return np.zeros(1, F32)
    step = F32(1.0) / F32(steps - 1)
    i = np.arange(steps)
    lo = (i.astype(F32) * step).astype(F32)
    hi = (1.0 - np.float64(step) * (steps - 1 - i).astype(np.float64)).astype(F32)
    return np.where(i < steps // 2, lo, hi).astype(F32)


def cumsum_f64acc(x: np.ndarray) -> np.ndarray:
    return np.cumsum(x.astype(np.float64), axis=-1).astype(F32)


def cumprod_f64acc(x: np.ndarray) -> np.ndarray:
    return np.cumprod(x.astype(np.float64), axis=-1).astype(F32)


def rowsum_f64(x: np.ndarray) -> np.ndarray:
    """Sequential float64 row sum rounded once to fp32 (see module docstring)."""
    acc = np.zeros(x.shape[:-1], np.float64)
    for j in range(x.shape[-1]):
        acc = acc + x[..., j].astype(np.float64)
    return acc.astype(F32)[..., None]


# ----------------------------------------------------------------------------------
# ray generation   (render_utils.py:23-41, 56-104; scene_utils.py:101-120)
# ----------------------------------------------------------------------------------
def get_ray_dirs(H: int, W: int, K: np.ndarray, flipped_yz: bool = False) -> np.ndarray:
    """render_utils.py:23-34 - pixel centres are the INTEGER coordinates (no +0.5)."""
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    xys = np.stack([xs, ys, np.ones_like(xs)], axis=-1).astype(F32)
    Kinv = np.linalg.inv(K.astype(F32)).astype(F32)
    dirs = (xys.reshape(-1, 3) @ Kinv.T).astype(F32).reshape(H, W, 3)
    return dirs * np.array([1, -1, -1], F32) if flipped_yz else dirs


def get_rays_c2w(dirs: np.ndarray, c2w: np.ndarray):
    """render_utils.py:37-41."""
    c2w = c2w.astype(F32)
    rays_d = (dirs.reshape(-1, 3) @ c2w[:3, :3].T).astype(F32).reshape(dirs.shape)
    rays_o = np.broadcast_to(c2w[:3, 3], rays_d.shape).astype(F32)
    norm = np.sqrt(np.sum(rays_d * rays_d, axis=-1, keepdims=True, dtype=F32))
    viewdirs = (rays_d / norm).astype(F32)
    return rays_o, rays_d, viewdirs


def rays_intersect_sphere(rays_o: np.ndarray, rays_d: np.ndarray, r: float = 1.0):
    """scene_utils.py:101-120 - forward intersection with the radius-r sphere.
    Raises AssertionError like the reference when a camera is outside the sphere."""
    odotd = np.sum(rays_o * rays_d, axis=1, dtype=F32)
    d_norm_sq = np.sum(rays_d * rays_d, axis=1, dtype=F32)
    o_norm_sq = np.sum(rays_o * rays_o, axis=1, dtype=F32)
    det = odotd * odotd + (F32(r * r) - o_norm_sq) * d_norm_sq
    assert np.all(det >= 0), "Not all your cameras are bounded by the unit sphere"
    return ((np.sqrt(det) - odotd) / d_norm_sq).astype(F32)


def prepare_rays_data(rays_o, rays_d, viewdirs, near, far, comp_radii: bool):
    """render_utils.py:81-104 on [H,W,3] inputs.  Radii: row y uses ||d[y]-d[y+1]||,
    the last row re-uses the pair (H-3, H-2) (``dx[-2:-1]``, :95)."""
    H, W, _ = rays_d.shape
    near = np.broadcast_to(np.asarray(near, F32), (H, W, 1)) if np.ndim(near) == 0 else near
    far = np.broadcast_to(np.asarray(far, F32), (H, W, 1)) if np.ndim(far) == 0 else far
    cols = [rays_o, rays_d, near.astype(F32), far.astype(F32)] + ([viewdirs] if viewdirs is not None else [])
    if comp_radii:
        diff = rays_d[:-1] - rays_d[1:]
        dx = np.sqrt(np.sum(diff * diff, axis=-1, dtype=F32))
        dx = np.concatenate([dx, dx[-2:-1]], axis=0)
        radii = (dx[..., None] * F32(2.0) / F32(np.sqrt(12.0))).astype(F32)
        cols.append(radii)
    return np.concatenate(cols, axis=-1).astype(F32)


def prepare_rays_from_pose(c2w, K, near=0.0, far=1.0, flipped_yz=False, comp_radii=False) -> np.ndarray:
    """render_utils.py:44-53: full-resolution rays, rays_d NOT normalised, constant near / far."""
    w, h = (K[:2, 2].astype(np.int32) * 2).tolist()
    rays_o, rays_d, viewdirs = get_rays_c2w(get_ray_dirs(h, w, K, flipped_yz), c2w)
    rays = prepare_rays_data(rays_o, rays_d, viewdirs, F32(near), F32(far), comp_radii)
    return np.ascontiguousarray(rays.reshape(-1, rays.shape[-1]))


def sample_nerf_rays(H, W, K, c2w, ds: int = 8, embed_type: str = "mip") -> np.ndarray:
    """render_utils.py:56-78 -> rays[n, 12] (mip) / [n, 11] (normal);
    columns o(3) d(3) near far viewdirs(3) [radius]; rays_d := viewdirs (:59)."""
    dirs = get_ray_dirs(H, W, K)
    rays_o, rays_d, viewdirs = get_rays_c2w(dirs, c2w)
    rays_d = viewdirs
    try:
        far = rays_intersect_sphere(rays_o.reshape(-1, 3), viewdirs.reshape(-1, 3), 1.0)
        far = far.reshape(H, W, 1)
    except AssertionError:
        far = np.ones((H, W, 1), F32)
    rays = prepare_rays_data(rays_o, rays_d, viewdirs, F32(0.01), far, embed_type == "mip")
    rays = rays[ds // 2 :: ds, ds // 2 :: ds]
    return np.ascontiguousarray(rays.reshape(-1, rays.shape[-1]))


# ----------------------------------------------------------------------------------
# mip-NeRF sampling   (render_utils.py:326-450, 453-615)
# ----------------------------------------------------------------------------------
def conical_frustum_to_gaussian(d, t0, t1, base_radius):
    """render_utils.py:350-380 (stable=True) + lift_gaussian :326-339 (diag=True).
    d [n,3], t0/t1 [n,S], base_radius [n,1] -> mean [n,S,3] (origin NOT added), cov_diag."""
    mu = (t0 + t1) / F32(2)
    hw = (t1 - t0) / F32(2)
    eps = F32(EPS32)
    mu2 = mu * mu
    hw2 = hw * hw
    hw4 = np.power(hw, F32(4.0)).astype(F32)
    denom = np.maximum(eps, F32(3) * mu2 + hw2)
    t_mean = mu + (F32(2) * mu * hw2) / denom
    t_var = hw2 / F32(3) - F32(4 / 15) * ((hw4 * (F32(12) * mu2 - hw2)) / (denom * denom))
    r_var = (base_radius * base_radius) * (mu2 / F32(4) + F32(5 / 12) * hw2 - F32(4 / 15) * hw4 / denom)
    mean = d[:, None, :] * t_mean[..., None]
    d_mag_sq = np.maximum(F32(1e-10), np.sum(d * d, axis=-1, keepdims=True, dtype=F32))
    d_outer = d * d
    null_outer = F32(1) - d_outer / d_mag_sq
    cov = t_var[..., None] * d_outer[:, None, :] + r_var[..., None] * null_outer[:, None, :]
    return mean.astype(F32), cov.astype(F32)


def cast_rays(t_vals, origins, directions, radii):
    """render_utils.py:383-402."""
    mean, cov = conical_frustum_to_gaussian(directions, t_vals[:, :-1], t_vals[:, 1:], radii)
    return (mean + origins[:, None, :]).astype(F32), cov


def sample_gaus_along_rays(rays: np.ndarray, S: int, t_rand: Optional[np.ndarray]):
    """render_utils.py:405-450.  ``t_rand`` [n,S+1] in [0,1) is the torch.rand draw of
    :444 (None -> randomized=False branch)."""
    o, d = rays[:, 0:3], rays[:, 3:6]
    near, far, radii = rays[:, 6:7], rays[:, 7:8], rays[:, 11:12]
    tau = torch_linspace01(S + 1)[None, :]
    t = near * (F32(1) - tau) + far * tau
    if t_rand is not None:
        mids = F32(0.5) * (t[:, 1:] + t[:, :-1])
        upper = np.concatenate([mids, t[:, -1:]], -1)
        lower = np.concatenate([t[:, :1], mids], -1)
        t = lower + (upper - lower) * t_rand.astype(F32)
    t = np.ascontiguousarray(t.astype(F32))
    mean, cov = cast_rays(t, o, d, radii)
    return t, mean, cov


def mip_resample_u(S1: int, u_jitter: Optional[np.ndarray], n: int) -> np.ndarray:
    """render_utils.py:477-496.  Note the reference's ``u + u + jitter`` (:480-486):
    the stratified offsets are DOUBLED, then clamped to 1-eps (:488-490)."""
    if u_jitter is not None:
        s = 1.0 / S1
        u = (np.arange(S1).astype(F32) * F32(s))[None, :]
        u = (u + u) + u_jitter.astype(F32)
        return np.minimum(u, F32(1.0 - EPS32)).astype(F32)
    # torch.linspace(0, 1-eps, S1): same two-sided formula with end = 1-eps
    end = F32(1.0 - EPS32)
    step = end / F32(S1 - 1)
    i = np.arange(S1)
    lo = (i.astype(F32) * step).astype(F32)
    hi = (np.float64(end) - np.float64(step) * (S1 - 1 - i)).astype(F32)
    return np.broadcast_to(np.where(i < S1 // 2, lo, hi).astype(F32), (n, S1)).copy()


def mip_blur_weights(weights: np.ndarray, resample_padding: float) -> np.ndarray:
    """render_utils.py:583-590: edge-pad, max of neighbours, mean of neighbours, +pad."""
    wp = np.concatenate([weights[:, :1], weights, weights[:, -1:]], -1)
    wmax = np.maximum(wp[:, :-1], wp[:, 1:])
    return (F32(0.5) * (wmax[:, :-1] + wmax[:, 1:]) + F32(resample_padding)).astype(F32)


def mip_cdf(weights: np.ndarray) -> np.ndarray:
    """render_utils.py:456-474: [n,S] blurred weights -> cdf [n,S+1] with exact 0 / 1 ends."""
    eps = F32(1e-5)
    S = weights.shape[-1]
    wsum = rowsum_f64(weights)
    padding = np.maximum(F32(0), eps - wsum)
    weights = weights + padding / F32(S)
    wsum = wsum + padding
    pdf = (weights / wsum).astype(F32)
    cdf = np.minimum(F32(1), cumsum_f64acc(pdf[:, :-1]))
    n = weights.shape[0]
    return np.concatenate([np.zeros((n, 1), F32), cdf, np.ones((n, 1), F32)], -1)


def invert_cdf_mip(bins: np.ndarray, cdf: np.ndarray, u: np.ndarray):
    """render_utils.py:498-552.  The reference's O(S^2) mask / where / max / min is,
    for sorted ``bins`` and non-decreasing ``cdf``, exactly
    ``idx = searchsorted(cdf, u, right=True) - 1`` (see _find_interval_bruteforce).
    Returns (samples [n,S1], idx int32 [n,S1])."""
    n, S1 = u.shape
    idx = np.empty((n, S1), np.int64)
    for r in range(n):
        idx[r] = np.searchsorted(cdf[r], u[r], side="right") - 1
    idx = np.clip(idx, 0, cdf.shape[1] - 2)
    x0 = np.take_along_axis(bins, idx, 1)
    x1 = np.take_along_axis(bins, idx + 1, 1)
    y0 = np.take_along_axis(cdf, idx, 1)
    y1 = np.take_along_axis(cdf, idx + 1, 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (u - y0) / (y1 - y0)
    t = np.nan_to_num(t, nan=0.0, posinf=np.finfo(F32).max, neginf=np.finfo(F32).min)
    t = np.clip(t, F32(0), F32(1)).astype(F32)
    return (x0 + t * (x1 - x0)).astype(F32), idx.astype(np.int32)


def _find_interval_bruteforce(bins, cdf, u):
    """Literal restatement of find_interval (render_utils.py:500-550) for small inputs;
    used by the tests to pin the searchsorted formulation above."""
    mask = u[:, None, :] >= cdf[:, :, None]

    def sel(x):
        x0 = np.max(np.where(mask, x[:, :, None], x[:, :1, None]), axis=-2)
        x1 = np.min(np.where(~mask, x[:, :, None], x[:, -1:, None]), axis=-2)
        return x0, x1

    x0, x1 = sel(bins)
    y0, y1 = sel(cdf)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (u - y0) / (y1 - y0)
    t = np.clip(np.nan_to_num(t, nan=0.0, posinf=np.finfo(F32).max, neginf=np.finfo(F32).min), 0, 1)
    return (x0 + t.astype(F32) * (x1 - x0)).astype(F32)


def resample_gaus_along_rays(rays, t_vals, weights, u_jitter, resample_padding=0.01):
    """render_utils.py:555-615.  The fine pass draws S+1 = t_vals.shape[-1] samples and
    ignores ``fine_nerf.num_pts`` (:592-597).  Returns new_t, mean, cov, idx."""
    o, d, radii = rays[:, 0:3], rays[:, 3:6], rays[:, 11:12]
    w = mip_blur_weights(weights.astype(F32), resample_padding)
    cdf = mip_cdf(w)
    u = mip_resample_u(t_vals.shape[-1], u_jitter, t_vals.shape[0])
    new_t, idx = invert_cdf_mip(t_vals, cdf, u)
    mean, cov = cast_rays(new_t, o, d, radii)
    return new_t, mean, cov, idx


# ----------------------------------------------------------------------------------
# classic NeRF sampling   (render_utils.py:107-148, 233-264)
# ----------------------------------------------------------------------------------
def sample_pts_init(rays, S: int, use_disp: bool = False, t_rand: Optional[np.ndarray] = None):
    """render_utils.py:107-129; ``t_rand`` [n,S] is the torch.rand_like draw of the perturb=True (training) branch
    (:118-126), None -> perturb=False (inference)."""
    o, d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8]
    tau = torch_linspace01(S)[None, :]
    if use_disp:
        z = F32(1.0) / (F32(1.0) / near * (F32(1) - tau) + F32(1.0) / far * tau)
    else:
        z = near * (F32(1) - tau) + far * tau
    z = np.broadcast_to(z.astype(F32), (rays.shape[0], S))
    if t_rand is not None:
        mids = F32(0.5) * (z[:, :-1] + z[:, 1:])
        upper = np.concatenate([mids, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mids], -1)
        z = lower + (upper - lower) * t_rand.astype(F32)
    z = np.ascontiguousarray(z.astype(F32))
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    return pts.astype(F32), z


def sample_pdf_det(bins, weights, N: int, u: Optional[np.ndarray] = None):
    """render_utils.py:233-264.  bins [n,B], weights [n,B-1] -> samples [n,N], inds int32 [n,N] (the raw searchsorted
    result of :251).  ``u`` [n,N]: the torch.rand draw of det=False (training); None -> det=True (linspace)."""
    w = (weights + F32(1e-5)).astype(F32)
    pdf = (w / rowsum_f64(w)).astype(F32)
    cdf = cumsum_f64acc(pdf)
    n = w.shape[0]
    cdf = np.concatenate([np.zeros((n, 1), F32), cdf], -1)
    u = np.broadcast_to(torch_linspace01(N), (n, N)) if u is None else np.ascontiguousarray(u, F32)
    inds = np.empty((n, N), np.int64)
    for r in range(n):
        inds[r] = np.searchsorted(cdf[r], u[r], side="right")
    below = np.maximum(0, inds - 1)
    above = np.minimum(cdf.shape[-1] - 1, inds)
    cb, ca = np.take_along_axis(cdf, below, 1), np.take_along_axis(cdf, above, 1)
    bb, ba = np.take_along_axis(bins, below, 1), np.take_along_axis(bins, above, 1)
    denom = ca - cb
    denom = np.where(denom < F32(1e-5), F32(1), denom)
    t = (u - cb) / denom
    return (bb + t * (ba - bb)).astype(F32), inds.astype(np.int32)


def sample_pts_fine(rays, z_vals, weights, N: int, u: Optional[np.ndarray] = None):
    """render_utils.py:132-148: 63 mid-point bins, weights[1:-1], sort(cat(z, samples))."""
    o, d = rays[:, 0:3], rays[:, 3:6]
    z_mid = F32(0.5) * (z_vals[:, 1:] + z_vals[:, :-1])
    samples, inds = sample_pdf_det(z_mid, weights[:, 1:-1], N, u)
    z = np.sort(np.concatenate([z_vals, samples], -1), axis=-1).astype(F32)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    return pts.astype(F32), z, inds


# ----------------------------------------------------------------------------------
# encodings   (embedding.py:13-84)
# ----------------------------------------------------------------------------------
HALF_PI32 = F32(0.5 * np.pi)


def ipe_encode(mean: np.ndarray, var: np.ndarray, num_freqs: int) -> np.ndarray:
    """PositionalEncodingMIP.forward with y=var, first output (embedding.py:66-79,
    renderer.py:156).  Layout: [sin block (freq-major, xyz-minor) | 'cos' block] where the
    cos block is sin(x + pi/2) evaluated in fp32; damping exp(-0.5 * var * 4^k)."""
    scales = (2.0 ** np.arange(num_freqs)).astype(F32)
    x = (mean[:, None, :] * scales[None, :, None]).reshape(mean.shape[0], -1)
    x = np.concatenate([x, x + HALF_PI32], -1).astype(F32)
    y = (var[:, None, :] * (scales * scales)[None, :, None]).reshape(var.shape[0], -1)
    y = np.concatenate([y, y], -1).astype(F32)
    return (np.exp(F32(-0.5) * y) * np.sin(x)).astype(F32)


def pe_dirs_encode(dirs: np.ndarray, num_freqs: int) -> np.ndarray:
    """PositionalEncodingMIP.forward with y=None (embedding.py:80-84): sin block, 'cos'
    block, then the raw input APPENDED."""
    scales = (2.0 ** np.arange(num_freqs)).astype(F32)
    x = (dirs[:, None, :] * scales[None, :, None]).reshape(dirs.shape[0], -1)
    x = np.concatenate([x, x + HALF_PI32], -1).astype(F32)
    return np.concatenate([np.sin(x), dirs], -1).astype(F32)


def fourier_encode(x: np.ndarray, num_freqs: int) -> np.ndarray:
    """FourierEmbedding.forward (embedding.py:35-46): [x, sin(f0 x), cos(f0 x), sin(f1 x), ...]."""
    out = [x.astype(F32)]
    for k in range(num_freqs):
        f = F32(2.0**k)
        out += [np.sin(f * x), np.cos(f * x)]
    return np.concatenate(out, -1).astype(F32)


# ----------------------------------------------------------------------------------
# NeRF MLP   (models/nerf.py:94-144)
# ----------------------------------------------------------------------------------
def _linear(x, W, b):
    return (x @ W.T + b).astype(F32)


def mlp_forward(wts: Dict[str, np.ndarray], x: np.ndarray, cfg: RenderConfig, stop_layer: Optional[int] = None):
    """NeRF.forward with use_viewdirs=True.  ``wts`` uses the reference state_dict key
    names without the ``nerf_coarse.`` / ``nerf_fine.`` prefix.  Returns
    (raw [P,4] = [sigmoid rgb, raw sigma], feat [P,hid])."""
    xd, dd = cfg.xyz_dim, cfg.dirs_dim
    in_pts, in_views, in_app = x[:, :xd], x[:, xd : xd + dd], x[:, xd + dd :]
    h = in_pts
    out_feat = None
    stop_layer = cfg.stop_layer if stop_layer is None else stop_layer
    for i in range(cfg.layer_num):
        h = np.maximum(_linear(h, wts[f"pts_linears.{i}.weight"], wts[f"pts_linears.{i}.bias"]), F32(0))
        if i == stop_layer:
            out_feat = h
        if i in cfg.skips:
            h = np.concatenate([in_pts, h], -1)
    pt_feat = h
    alpha = _linear(h, wts["alpha_linear.weight"], wts["alpha_linear.bias"])
    feature = _linear(h, wts["feature_linear.weight"], wts["feature_linear.bias"])
    h_rgb = np.concatenate([feature, in_views, in_app], -1)
    h_rgb = np.maximum(_linear(h_rgb, wts["views_linears.0.weight"], wts["views_linears.0.bias"]), F32(0))
    rgb = _linear(h_rgb, wts["rgb_linear.weight"], wts["rgb_linear.bias"])
    rgb = (F32(1) / (F32(1) + np.exp(-rgb))).astype(F32)
    raw = np.concatenate([rgb, alpha], -1).astype(F32)
    feat = out_feat if stop_layer >= 0 else pt_feat
    return raw, feat


def round_bf16(x: np.ndarray) -> np.ndarray:
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.ascontiguousarray(x, F32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(F32)


def round_fp16(x: np.ndarray) -> np.ndarray:
    """round-to-nearest-even to IEEE half, saturating at +-65504, returned as float32"""
    return np.clip(np.ascontiguousarray(x, F32), -65504, 65504).astype(np.float16).astype(F32)


def mlp_forward_rounded(wts: Dict[str, np.ndarray], x: np.ndarray, cfg: RenderConfig, stop_layer: Optional[int] = None, rnd=round_bf16):
    """NeRF.forward with 16-bit tensor-core operands, as ANY implementation that feeds the 8 x 256 trunk to bf16 / fp16
    matrix units computes it: weights and layer inputs rounded by ``rnd`` (round_bf16 | round_fp16), products accumulated in
    fp32, biases / the sigma head / the dirs part of views_linears / rgb in fp32.  Used to measure how far 16-bit operands
    alone move the reference's outputs (the error envelope the CUDA path is judged against), not as a reference itself."""
    xd = cfg.xyz_dim
    stop_layer = cfg.stop_layer if stop_layer is None else stop_layer
    enc = rnd(x[:, :xd])
    side = x[:, xd:]
    W = lambda k: rnd(wts[k])
    h, feat = enc, None
    for i in range(cfg.layer_num):
        w = W(f"pts_linears.{i}.weight")
        pre = (enc @ w[:, :xd].T + h @ w[:, xd:].T) if (i - 1) in cfg.skips else h @ w.T
        act = np.maximum(pre + wts[f"pts_linears.{i}.bias"], 0).astype(F32)
        h = rnd(act)
        if i == stop_layer:
            feat = h
    if stop_layer < 0:
        feat = h
    alpha = h @ wts["alpha_linear.weight"].astype(F32).T + wts["alpha_linear.bias"]
    wv = wts["views_linears.0.weight"].astype(np.float64)
    wfold = (wv[:, :cfg.hid_dim] @ wts["feature_linear.weight"].astype(np.float64)).astype(F32)
    bfold = (wts["views_linears.0.bias"] + wv[:, :cfg.hid_dim] @ wts["feature_linear.bias"].astype(np.float64)).astype(F32)
    hv = np.maximum(h @ rnd(wfold).T + (side @ wts["views_linears.0.weight"][:, cfg.hid_dim:].T + bfold), 0).astype(F32)
    rgb = F32(1) / (F32(1) + np.exp(-(hv @ wts["rgb_linear.weight"].T + wts["rgb_linear.bias"])))
    return np.concatenate([rgb, alpha], -1).astype(F32), feat.astype(F32)


def encode_points(cfg: RenderConfig, pts, var, viewdirs, app_emb, S: int) -> np.ndarray:
    """renderer.py:125-164: per-sample [xyz enc | dirs enc | appearance] rows."""
    n = pts.shape[0]
    pf = pts.reshape(n * S, 3)
    if cfg.embed_type == "mip":
        x = ipe_encode(pf, var.reshape(n * S, 3), cfg.xyz_num_freqs)
    else:
        x = fourier_encode(pf, cfg.xyz_num_freqs)
    cols = [x]
    if cfg.use_viewdirs:
        enc = pe_dirs_encode if cfg.embed_type == "mip" else fourier_encode
        de = enc(viewdirs.astype(F32), cfg.dirs_num_freqs)
        cols.append(np.repeat(de, S, axis=0))
    if app_emb is not None:
        cols.append(np.repeat(app_emb.astype(F32), S, axis=0))
    return np.concatenate(cols, -1).astype(F32)


# ----------------------------------------------------------------------------------
# volume rendering + compositing   (render_utils.py:176-230, renderer.py:250-281)
# ----------------------------------------------------------------------------------
def t_to_s(z: np.ndarray) -> np.ndarray:
    """render_utils.py:618-636 as renderer.py:283-285 calls it: near / far are the min / max over the WHOLE batch, and
    g(x) = 1 / (x + 1e-6) adds its epsilon IN PLACE (:634) - ``near`` is passed to g twice, so the second call (in the
    denominator) sees near + 2e-6: s = (1/(t+e) - 1/(near+e)) / (1/(far+e) - 1/(near+e+e))."""
    eps = F32(1e-6)
    near1 = z.min() + eps
    near2 = near1 + eps
    return ((F32(1) / (z + eps) - F32(1) / near1) / (F32(1) / (z.max() + eps) - F32(1) / near2)).astype(F32)


def volume_render(raw: np.ndarray, z_vals: np.ndarray, rays_d: np.ndarray, mip: bool, white_bg=False, noise: Optional[np.ndarray] = None):
    """volume_render_radiance_field; ``noise`` [n,S] = randn * noise_std of the training mode (None: inference).  raw [n,S,4]."""
    rgb = raw[..., :3]
    density = np.maximum(raw[..., 3] + (F32(0) if noise is None else noise.astype(F32)), F32(0))
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    if mip:
        z_mids = F32(0.5) * (z_vals[:, :-1] + z_vals[:, 1:])
    else:
        dists = np.concatenate([dists, np.full_like(dists[:, :1], 1e10)], -1)
        z_mids = z_vals
    dnorm = np.sqrt(np.sum(rays_d * rays_d, axis=-1, keepdims=True, dtype=F32))
    dists = (dists * dnorm).astype(F32)
    alpha = (F32(1.0) - np.exp(-density * dists)).astype(F32)
    trans_in = np.concatenate([np.ones_like(alpha[:, :1]), F32(1.0) - alpha + F32(1e-10)], -1)
    T = cumprod_f64acc(trans_in)[:, :-1]
    weights = (alpha * T).astype(F32)
    rgb_map = np.sum(weights[..., None] * rgb, axis=-2, dtype=F32)
    depth = np.sum(weights * z_mids, axis=-1, dtype=F32)
    acc = np.sum(weights, axis=-1, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        disp = F32(1.0) / np.maximum(F32(1e-10), depth / acc)
    if white_bg:
        rgb_map = rgb_map + (F32(1.0) - acc[:, None])
    return rgb_map.astype(F32), disp.astype(F32), acc, weights, depth


def composite(weights, feats, pts, feat_comb: str):
    """renderer.py:257-265, 276-281.  feats [n,S,C], pts [n,S,3]."""
    if feat_comb == "max":
        ids = np.argmax(weights, axis=-1)
        ar = np.arange(weights.shape[0])
        return feats[ar, ids, :].astype(F32), pts[ar, ids, :].astype(F32)
    f = np.einsum("ns,nsc->nc", weights.astype(np.float64), feats.astype(np.float64)).astype(F32)
    p = np.sum(weights[..., None] * pts, axis=-2, dtype=F32)
    return f, p


def unnormalize_pts(pts: np.ndarray, unnorm: np.ndarray) -> np.ndarray:
    """utils/geometry.py:76-85: homogeneous 4x4 affine, first three rows."""
    ph = np.concatenate([pts, np.ones_like(pts[:, :1])], -1).astype(F32)
    return (ph @ unnorm.astype(F32).T)[:, :3].astype(F32)


# ----------------------------------------------------------------------------------
# the whole path   (renderer.py:182-295)
# ----------------------------------------------------------------------------------
def split_state_dict(sd: Dict[str, np.ndarray], prefix: str) -> Dict[str, np.ndarray]:
    return {k[len(prefix) :]: np.asarray(v, F32) for k, v in sd.items() if k.startswith(prefix)}


def render_rays(
    cfg: RenderConfig,
    sd: Dict[str, np.ndarray],
    rays: np.ndarray,
    t_rand: Optional[np.ndarray] = None,
    u_jitter: Optional[np.ndarray] = None,
    ray_id: Optional[np.ndarray] = None,
    keep_intermediates: bool = False,
    chunk: int = 4096,
    validation: bool = True,
    noise_coarse: Optional[np.ndarray] = None,
    noise_fine: Optional[np.ndarray] = None,
    mlp=None,
) -> Dict[str, np.ndarray]:
    """NerfRenderer.render_rays (renderer.py:182-295), chunked over rays only to bound memory (rays are independent).
    Training mode (validation=False, one chunk only because s_fine normalises over the batch): ``t_rand`` / ``u_jitter``
    are then also the classic mode's perturbation [n,Sc] and sample_pdf draws [n,Nf]; ``noise_*`` the scaled sigma noise;
    outputs drop pts_* and gain s_fine / weights_fine (:276-286)."""
    rays = np.ascontiguousarray(rays, F32)
    n = rays.shape[0]
    if not validation:
        chunk = max(chunk, n)
    if n > chunk:
        parts = []
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            parts.append(
                render_rays(
                    cfg, sd, rays[s:e],
                    None if t_rand is None else t_rand[s:e],
                    None if u_jitter is None else u_jitter[s:e],
                    None if ray_id is None else ray_id[s:e],
                    keep_intermediates, chunk, mlp=mlp,
                )
            )
        return {k: np.concatenate([p[k] for p in parts], 0) for k in parts[0]}

    mip = cfg.embed_type == "mip"
    rays_d = rays[:, 3:6]
    viewdirs = rays[:, 8:11] if rays.shape[1] >= 11 else rays_d
    app = None
    if cfg.app_dim > 0:
        rid = np.ones(n, np.int64) if ray_id is None else ray_id  # renderer.py:298-299
        app = np.asarray(sd["embedding_a.weight"], F32)[rid]
    out: Dict[str, np.ndarray] = {}
    z_vals = weights = None
    for key in ("coarse", "fine"):
        w = split_state_dict(sd, f"nerf_{key}.")
        if mip:
            if key == "coarse":
                z_vals, mean, var = sample_gaus_along_rays(rays, cfg.num_pts_coarse, t_rand)
            else:
                z_vals, mean, var, idx = resample_gaus_along_rays(
                    rays, z_vals, weights, u_jitter, cfg.resample_padding
                )
                if keep_intermediates:
                    out["idx_fine"] = idx
            if cfg.mip_var_scale > 0:
                var = (F32(cfg.mip_var_scale) * var).astype(F32)
            pts = mean
            S = z_vals.shape[1] - 1
        else:
            if key == "coarse":
                pts, z_vals = sample_pts_init(rays, cfg.num_pts_coarse, cfg.use_disp, t_rand)
            else:
                pts, z_vals, idx = sample_pts_fine(rays, z_vals, weights, cfg.num_pts_fine, u_jitter)
                if keep_intermediates:
                    out["idx_fine"] = idx
            var = None
            S = z_vals.shape[1]
        x = encode_points(cfg, pts, var, viewdirs, app, S)
        raw, feat = (mlp or mlp_forward)(w, x, cfg, cfg.stop_layer if key == "fine" else cfg.stop_layer_coarse)
        raw = raw.reshape(n, S, 4)
        feat = feat.reshape(n, S, -1)
        rgb_map, disp, acc, weights, depth = volume_render(raw, z_vals, rays_d, mip, cfg.white_bg, noise_coarse if key == "coarse" else noise_fine)
        f, p = composite(weights, feat, pts, cfg.feat_comb)
        if cfg.ret_pfeat:
            out[f"feat_{key}"] = f
        if validation:
            out[f"pts_{key}"] = p
        elif key == "fine":
            out["s_fine"] = t_to_s(z_vals)
            out["weights_fine"] = weights
        out[f"rgb_{key}"] = rgb_map
        out[f"depth_{key}"] = depth
        if keep_intermediates:
            out[f"z_{key}"] = z_vals
            out[f"weights_{key}"] = weights
            out[f"raw_{key}"] = raw
            out[f"acc_{key}"] = acc
    return out
