"""Generate tests/golden/*.npz by running the UNMODIFIED reference renderer (CPU, fp32).

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

The reference imports two off-path packages (pycolmap, imageio) at module import; they are
stubbed with empty modules exactly as the survey did.  Weights come from
oracle/weights.py (seeded hash), so fixtures store only inputs, random draws, intermediates
and outputs.  Intermediates are captured by wrapping the two module-level functions that
``renderer.render_rays`` calls (sample_smth_along_rays / volume_render_radiance_field).
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace as NS

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True

REF = "/root/reference"


def import_reference():
    import torch  # noqa: F401

    sys.path.insert(0, REF)
    for m in ("pycolmap", "imageio"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import nerfmatch.nerf.renderer as r

    return r


def ref_config(embed="mip", nc=64, nf=128, app=False, white_bg=False):
    # mirrors configs/nerf/nerf_7scenes_mip_sfm.yaml:24-55
    emb = NS(xyz_num_freqs=15, dirs_num_freqs=4, appearance_embed=app)
    if embed is not None:
        emb.type = embed
    return NS(
        data=NS(img_wh=[160, 120]),
        render=NS(chunksize=16384, use_disp=False, perturb=True, noise_std=1.0, white_bg=white_bg, use_viewdirs=True),
        embedding=emb,
        coarse_nerf=NS(method="NeRF", layer_num=8, hid_dim=256, output_dim=4, skips=[4], num_pts=nc),
        fine_nerf=NS(method="NeRF", layer_num=8, hid_dim=256, output_dim=4, skips=[4], num_pts=nf),
        loss=NS(),
    )


CASES = {
    # sigma_target = (mean, std) the raw density is calibrated to on this view (see calibrate_sigma)
    "mip64": dict(embed="mip", nc=64, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False, sigma_target=(2.0, 1.0), ds=48, seed=0),
    "mip128_peaky": dict(embed="mip", nc=128, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False, sigma_target=(5.0, 40.0), ds=60, seed=1),
    "normal_64_128": dict(embed="normal", nc=64, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False, sigma_target=(2.0, 1.0), ds=48, seed=2),
    "normal_64_128_peaky": dict(embed="normal", nc=64, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False, sigma_target=(5.0, 40.0), ds=60, seed=3),
    "mip64_app_max_white": dict(embed="mip", nc=64, nf=128, app=True, stop_layer=-1, feat_comb="max", white_bg=True, sigma_target=(5.0, 40.0), ds=60, seed=4),
    "normal_app_stop5": dict(embed="normal", nc=64, nf=128, app=True, stop_layer=5, feat_comb="lin", white_bg=False, sigma_target=(5.0, 40.0), ds=80, seed=5),
}


def calibrate_sigma(c, rays):
    """Default-init densities are all slightly negative (every weight would be 0): measure the raw sigma
    of each net on this view's coarse samples and return the (gain, bias) pairs that map it to the
    case's target mean / std.  Rounded to 4 significant digits so the case dict reproduces them."""
    from oracle import nerf_oracle as O
    from oracle.weights import make_state_dict

    app = 16 if c["app"] else 0
    sd = make_state_dict(c["seed"], embed_type=c["embed"], app_dim=app, num_frames=7 if c["app"] else 0)
    cfg = O.RenderConfig(embed_type=c["embed"], num_pts_coarse=c["nc"], app_dim=app, stop_layer=c["stop_layer"])
    n = rays.shape[0]
    if c["embed"] == "mip":
        _, pts, var = O.sample_gaus_along_rays(rays, c["nc"], None)
    else:
        pts, _ = O.sample_pts_init(rays, c["nc"])
        var = None
    appv = sd["embedding_a.weight"][np.arange(n) % 7] if app else None
    x = O.encode_points(cfg, pts, var, rays[:, 8:11], appv, c["nc"])
    gains, biases = [], []
    for net in ("nerf_coarse.", "nerf_fine."):
        raw, _ = O.mlp_forward(O.split_state_dict(sd, net), x, cfg)
        m, s = float(raw[:, 3].mean()), float(raw[:, 3].std())
        g = float(f"{c['sigma_target'][1] / s:.4g}")
        gains.append(g)
        biases.append(float(f"{c['sigma_target'][0] - g * m:.4g}"))
    return tuple(gains), tuple(biases)


def run_case(r, name, c):
    import torch

    from oracle.weights import K_7SCENES, make_pose, make_state_dict, make_unnorm_scene

    c = dict(c)
    rays0 = r.sample_nerf_rays(480, 640, torch.from_numpy(K_7SCENES), torch.from_numpy(make_pose(c["seed"])), ds=c["ds"], embed_type=c["embed"])
    c["sigma_gain"], c["sigma_bias"] = calibrate_sigma(c, rays0.numpy())

    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    num_frames = 7 if c["app"] else None
    ren = r.NerfRenderer(
        ref_config(c["embed"], c["nc"], c["nf"], c["app"], c["white_bg"]),
        num_frames=num_frames, training=False, stop_layer=c["stop_layer"],
    ).eval()
    sd = make_state_dict(
        c["seed"], embed_type=c["embed"], app_dim=16 if c["app"] else 0, num_frames=num_frames or 0,
        sigma_gain=c["sigma_gain"], sigma_bias=c["sigma_bias"],
    )
    missing = ren.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    ren.ret_pfeat = True
    ren.feat_comb = c["feat_comb"]

    H, W = 480, 640
    K = torch.from_numpy(K_7SCENES)
    c2w = torch.from_numpy(make_pose(c["seed"]))
    rays = r.sample_nerf_rays(H, W, K, c2w, ds=c["ds"], embed_type=c["embed"])
    n = rays.shape[0]
    h, w = len(range(c["ds"] // 2, H, c["ds"])), len(range(c["ds"] // 2, W, c["ds"]))
    assert h * w == n
    ray_id = None
    if c["app"]:
        ray_id = torch.from_numpy((np.arange(n) % 7).astype(np.int64))

    # capture intermediates
    cap = {"z": [], "w": [], "raw": []}
    orig_sample, orig_vol = r.sample_smth_along_rays, r.volume_render_radiance_field

    def sample_wrap(*a, **k):
        out = orig_sample(*a, **k)
        cap["z"].append(out[1].clone().numpy())
        return out

    def vol_wrap(raw, *a, **k):
        out = orig_vol(raw, *a, **k)
        cap["raw"].append(raw.clone().numpy())
        cap["w"].append(out[3].clone().numpy())
        return out

    r.sample_smth_along_rays, r.volume_render_radiance_field = sample_wrap, vol_wrap
    try:
        torch.manual_seed(7)
        preds = ren.predict(rays, w, h, ray_id=ray_id)
    finally:
        r.sample_smth_along_rays, r.volume_render_radiance_field = orig_sample, orig_vol

    out = {
        "rays": rays.numpy(), "hw": np.array([h, w]), "H": H, "W": W, "ds": c["ds"],
        "K": K.numpy(), "c2w": c2w.numpy(),
    }
    if ray_id is not None:
        out["ray_id"] = ray_id.numpy()
    if c["embed"] == "mip":
        # replay the two draws the reference consumed (render_utils.py:444, 483-485)
        S1 = c["nc"] + 1
        torch.manual_seed(7)
        t_rand = torch.rand(n, S1)
        s = 1 / S1
        u_jit = torch.empty(n, S1).uniform_(to=(s - torch.finfo(torch.float32).eps))
        out["t_rand"], out["u_jitter"] = t_rand.numpy(), u_jit.numpy()
    for key in ("coarse", "fine"):
        i = 0 if key == "coarse" else 1
        out[f"z_{key}"] = cap["z"][i]
        out[f"weights_{key}"] = cap["w"][i]
        out[f"raw_{key}"] = cap["raw"][i]
        out[f"feat_{key}"] = preds[f"feat_{key}"].numpy()
        out[f"pts_{key}"] = preds[f"pts_{key}"].numpy()
        out[f"rgb_{key}"] = preds[f"rgb_{key}"].reshape(n, 3).numpy()
        out[f"depth_{key}"] = preds[f"depth_{key}"].reshape(n).numpy()
    # un-normalised points as render_novel_view returns them (renderer.py:327)
    un = make_unnorm_scene()
    from nerfmatch.utils.geometry import unnormaliz_pts

    out["unnorm_scene"] = un
    out["pt3d"] = unnormaliz_pts(preds["pts_fine"][None], torch.from_numpy(un)[None])[0].numpy()
    out["case"] = np.array(repr(c))
    return out


def raygen_cases(r):
    import torch

    from oracle.weights import K_7SCENES, K_CAMBRIDGE, make_pose

    out = {}
    for tag, (H, W, K, ds, embed) in {
        "7s_ds8_mip": (480, 640, K_7SCENES, 8, "mip"),
        "7s_ds8_normal": (480, 640, K_7SCENES, 8, "normal"),
        "small_ds1_mip": (24, 40, np.array([[30.0, 0, 20], [0, 31.0, 12], [0, 0, 1]], np.float32), 1, "mip"),
        "cam_ds16_mip": (480, 854, K_CAMBRIDGE, 16, "mip"),
    }.items():
        c2w = make_pose(11)
        rays = r.sample_nerf_rays(H, W, torch.from_numpy(K), torch.from_numpy(c2w), ds=ds, embed_type=embed)
        out[f"{tag}.rays"] = rays.numpy()
        out[f"{tag}.meta"] = np.array([H, W, ds, 1 if embed == "mip" else 0])
        out[f"{tag}.K"] = K
        out[f"{tag}.c2w"] = c2w
    return out


def stage_cases(r):
    """Direct calls of reference stage functions on synthetic inputs (peaky weights etc.)."""
    import torch
    from nerfmatch.nerf import render_utils as ru
    from nerfmatch.nerf.embedding import FourierEmbedding, PositionalEncodingMIP

    from oracle.weights import uniform01

    out = {}
    n, S = 96, 64
    # --- mip resampling on hand-made weights: near-delta, zeros, uniform, random
    t = np.sort(uniform01(3, 1, n * (S + 1)).reshape(n, S + 1) * 1.3 + 0.01, -1).astype(np.float32)
    w = uniform01(3, 2, n * S).reshape(n, S).astype(np.float32) ** 8
    w[0] = 0.0
    w[1] = 1.0 / S
    w[2] = 0.0
    w[2, 17] = 0.9
    w[3] = 0.0
    w[3, 0] = 0.7
    w[4] = 0.0
    w[4, -1] = 0.99
    origins = (uniform01(3, 3, n * 3).reshape(n, 3) * 0.4 - 0.2).astype(np.float32)
    d = uniform01(3, 4, n * 3).reshape(n, 3) - 0.5
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    radii = (uniform01(3, 5, n).reshape(n, 1) * 1e-3 + 5e-4).astype(np.float32)
    torch.manual_seed(5)
    s = 1 / (S + 1)
    u_jit = torch.empty(n, S + 1).uniform_(to=(s - torch.finfo(torch.float32).eps))
    for randomized in (True, False):
        torch.manual_seed(5)
        new_t, (mean, cov) = ru.resample_gaus_along_rays(
            torch.from_numpy(origins), torch.from_numpy(d), torch.from_numpy(radii), torch.from_numpy(t.copy()),
            torch.from_numpy(w.copy()), randomized=randomized, stop_grad=True, resample_padding=0.01, ray_shape="cone",
        )
        tag = "rand" if randomized else "det"
        out[f"mip_resample.{tag}.new_t"] = new_t.numpy()
        out[f"mip_resample.{tag}.mean"] = mean.numpy()
        out[f"mip_resample.{tag}.cov"] = cov.numpy()
    out["mip_resample.t"], out["mip_resample.w"], out["mip_resample.u_jitter"] = t, w, u_jit.numpy()
    # reference-side interval indices: the same torch ops as render_utils.py:456-490, 583-590 (so the
    # row sum has torch's own order), then the searchsorted form of find_interval.
    tw = torch.from_numpy(w.copy())
    wp = torch.cat([tw[..., :1], tw, tw[..., -1:]], dim=-1)
    wm = torch.maximum(wp[..., :-1], wp[..., 1:])
    wb = 0.5 * (wm[..., :-1] + wm[..., 1:]) + 0.01
    wsum = torch.sum(wb, dim=-1, keepdim=True)
    pad = torch.maximum(torch.zeros_like(wsum), 1e-5 - wsum)
    wb = wb + pad / wb.shape[-1]
    wsum = wsum + pad
    cdf = torch.minimum(torch.ones(()), torch.cumsum((wb / wsum)[..., :-1], dim=-1))
    cdf = torch.cat([torch.zeros(n, 1), cdf, torch.ones(n, 1)], dim=-1)
    uu = (torch.arange(S + 1) * s)[None, ...]
    uu = torch.minimum(uu + uu + u_jit, torch.full_like(u_jit, 1.0 - torch.finfo(torch.float32).eps))
    out["mip_resample.rand.idx_ref"] = (torch.searchsorted(cdf, uu.contiguous(), right=True) - 1).numpy().astype(np.int32)
    out["mip_resample.rand.cdf_ref"] = cdf.numpy()
    out["mip_resample.o"], out["mip_resample.d"], out["mip_resample.radii"] = origins, d, radii

    # --- classic sample_pdf + sort
    z = np.sort(uniform01(4, 1, n * S).reshape(n, S) * 1.2 + 0.01, -1).astype(np.float32)
    wn = uniform01(4, 2, n * S).reshape(n, S).astype(np.float32) ** 6
    wn[0] = 0
    wn[1] = 0
    wn[1, 30] = 0.8
    wn[2] = 0
    wn[2, 1] = 0.5
    wn[2, 62] = 0.5
    rays = np.concatenate([origins, d, np.full((n, 1), 0.01, np.float32), np.full((n, 1), 1.2, np.float32), d], -1).astype(np.float32)
    pts, z_new = ru.sample_pts_fine(torch.from_numpy(origins), torch.from_numpy(d), torch.from_numpy(z), torch.from_numpy(wn), num_pts=128, perturb=False)
    out["normal_resample.z"], out["normal_resample.w"], out["normal_resample.rays"] = z, wn, rays
    out["normal_resample.z_new"], out["normal_resample.pts"] = z_new.numpy(), pts.numpy()

    # --- encodings at large arguments
    x = (uniform01(6, 1, 200 * 3).reshape(200, 3) * 2 - 1).astype(np.float32)
    v = (uniform01(6, 2, 200 * 3).reshape(200, 3) * 1e-4).astype(np.float32)
    mipenc = PositionalEncodingMIP(15)
    out["enc.x"], out["enc.var"] = x, v
    out["enc.ipe"] = mipenc(torch.from_numpy(x), y=torch.from_numpy(v))[0].numpy()
    out["enc.pe_dirs"] = PositionalEncodingMIP(4)(torch.from_numpy(x)).numpy()
    out["enc.fourier15"] = FourierEmbedding(15)(torch.from_numpy(x)).numpy()
    out["enc.fourier4"] = FourierEmbedding(4)(torch.from_numpy(x)).numpy()

    # --- volume rendering (both modes, white bg) on synthetic raw
    raw = (uniform01(7, 1, n * S * 4).reshape(n, S, 4)).astype(np.float32)
    raw[..., 3] = (raw[..., 3] - 0.4) * 60.0
    raw[5, :, 3] = -1.0  # empty ray
    raw[6, :, 3] = 1e4  # opaque at first sample
    for mode, zz in (("mip", t), ("normal", z)):
        for wb in (False, True):
            res = ru.volume_render_radiance_field(
                torch.from_numpy(raw), torch.from_numpy(zz), torch.from_numpy(d), noise_std=0.0, white_bg=wb, embed_type=mode,
            )
            tag = f"volrender.{mode}.{'white' if wb else 'black'}"
            out[tag + ".rgb"], out[tag + ".disp"], out[tag + ".acc"] = res[0].numpy(), res[1].numpy(), res[2].numpy()
            out[tag + ".weights"], out[tag + ".depth"] = res[3].numpy(), res[4].numpy()
    out["volrender.raw"] = raw
    return out


def match_cases():
    """Reference coarse matching: NeRFMatcherCoarse.coarse_matching called unbound on a namespace that carries the two
    attributes __init__ would set (nerfmatch_coarse_trainer.py:69-83), then the reference's extract_mutual_matches.
    Heavy off-path imports (lightning, timm, kornia, ...) are stubbed."""
    import importlib
    from types import SimpleNamespace

    import torch

    from oracle.weights import uniform01

    class _Any(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            m = _Any(self.__name__ + "." + name)
            setattr(self, name, m)
            return m

        def __call__(self, *a, **k):
            return self

    for m in ("pytorch_lightning", "pytorch_lightning.callbacks", "pytorch_lightning.loggers", "pytorch_lightning.strategies",
              "timm", "kornia", "kornia.geometry", "kornia.geometry.subpix", "kornia.utils", "kornia.utils.grid", "transforms3d",
              "transforms3d.quaternions"):
        if m not in sys.modules:
            try:
                importlib.import_module(m)
            except Exception:
                sys.modules[m] = _Any(m)
    sys.modules["pytorch_lightning"].LightningModule = torch.nn.Module
    import nerfmatch.nerfmatch_coarse_trainer as ct
    from nerfmatch.modules.extract_matches import extract_mutual_matches

    out = {}
    m, n, d = 300, 350, 256
    pt = (uniform01(8, 1, n * d).reshape(n, d) - 0.5).astype(np.float32)
    perm = (uniform01(8, 2, m) * n).astype(np.int64)
    im = (pt[perm] + 0.35 * (uniform01(8, 3, m * d).reshape(m, d) - 0.5)).astype(np.float32)    # noisy copies -> real matches
    im_mask = uniform01(8, 4, m) > 0.1
    pt_mask = uniform01(8, 5, n) > 0.1
    out["im"], out["pt"], out["im_mask"], out["pt_mask"] = im, pt, im_mask, pt_mask
    for tag, temp, ttype, masks in (("mul10", 10.0, "mul", False), ("div01_masked", 0.1, "div", True)):
        t = torch.tensor(temp)
        fake = SimpleNamespace(
            norm_feat=lambda feat: (feat / (feat.norm(dim=-1, keepdim=True) + 1e-6)),
            apply_temperature=(lambda s: s / t) if ttype == "div" else (lambda s: s * t),
        )
        conf, _, _ = ct.NeRFMatcherCoarse.coarse_matching(
            fake, torch.from_numpy(im)[None], torch.from_numpy(pt)[None],
            im_mask=torch.from_numpy(im_mask)[None] if masks else None, pt_mask=torch.from_numpy(pt_mask)[None] if masks else None)
        out[f"{tag}.conf"] = conf[0].numpy()
        for mutual in (False, True):
            for thr in (0.0, 0.2):
                (b_ids, i_ids, j_ids), mconf, num = extract_mutual_matches(conf, mutual=mutual, threshold=thr)
                k = f"{tag}.{'mutual' if mutual else 'rowmax'}.thr{thr}"
                out[k + ".i"], out[k + ".j"], out[k + ".mconf"] = i_ids.numpy(), j_ids.numpy(), mconf.numpy()
    return out


def rayhelper_cases(r):
    """get_ray_dirs / get_rays_c2w / prepare_rays_data / prepare_rays_from_pose of the reference (render_utils.py:23-53, 81-104)."""
    import torch

    from oracle.weights import make_pose

    ru = sys.modules["nerfmatch.nerf.render_utils"]
    K = np.array([[30.0, 0, 20], [0, 31.0, 12], [0, 0, 1]], np.float32)      # -> w, h = 40, 24 in prepare_rays_from_pose
    H, W = 24, 40
    c2w = make_pose(12)
    out = {"K": K, "c2w": c2w, "hw": np.array([H, W])}
    for flip in (0, 1):
        dirs, xys = ru.get_ray_dirs(H, W, torch.from_numpy(K), flipped_yz=bool(flip), return_xys=True)
        out[f"dirs{flip}"] = dirs.numpy()
        out[f"xys{flip}"] = xys.numpy()
    o, d, v = ru.get_rays_c2w(torch.from_numpy(out["dirs0"]), torch.from_numpy(c2w))
    out["rays_o"], out["rays_d"], out["viewdirs"] = o.numpy().copy(), d.numpy(), v.numpy()
    far = torch.from_numpy(np.linspace(0.7, 1.9, H * W, dtype=np.float32).reshape(H, W, 1))
    out["far"] = far.numpy()
    out["prep_mip"] = ru.prepare_rays_data(o, d, v, 0.01, far, comp_radii=True).numpy()            # tensor far, radii
    out["prep_plain"] = ru.prepare_rays_data(o, v, v, 0.05, 2.5, comp_radii=False).numpy()          # scalars, rays_d := viewdirs
    out["prep_nodirs"] = ru.prepare_rays_data(o, d, None, 0.0, 1.0, flatten=False, comp_radii=True).numpy()
    for flip in (0, 1):
        out[f"pose{flip}"] = ru.prepare_rays_from_pose(c2w, torch.from_numpy(K), 0.2, 3.0, flipped_yz=bool(flip), comp_radii=True).numpy()
    out["pose_plain"] = ru.prepare_rays_from_pose(c2w, torch.from_numpy(K), 0.0, 1.0).numpy()
    return out


def inerf_cases(r):
    """One iNeRF refinement step's differentiable graph (nerfmatch_evaluator.py:236-285 ray generation from the pose,
    :336-424 coarse pass without grad, fine pass with grad, white-background volume rendering, MSE photometric loss),
    built from the reference's own functions under torch autograd; the fixture holds the loss and d loss / d pose.
    (The evaluator module itself needs pytorch_lightning / pycolmap / kornia, so its few glue lines are restated.)"""
    import torch

    from oracle.weights import K_7SCENES, make_pose, make_state_dict

    torch.set_grad_enabled(True)
    out = {}
    for tag, seed, gain, bias in (("smooth", 21, 60.0, 4.0), ("peaky", 22, 400.0, 20.0)):
        torch.manual_seed(0)
        ren = r.NerfRenderer(ref_config("mip", 128, 128, False, True), num_frames=None, training=False, stop_layer=3).eval()
        sd = make_state_dict(seed, embed_type="mip", sigma_gain=gain, sigma_bias=(bias, bias))
        ren.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        for prm in ren.parameters():
            prm.requires_grad_(False)
        H, W, ds = 480, 640, 40
        K = torch.from_numpy(K_7SCENES)
        cam_pose = torch.from_numpy(make_pose(seed).astype(np.float32))[None].clone().requires_grad_(True)
        # gen_rays (nerfmatch_evaluator.py:236-285)
        ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        xys = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).float()
        dirs = xys @ torch.linalg.inv(K).T
        cam_centers = cam_pose[:, None, None, :3, 3].expand(-1, H, W, -1)[0]
        cam_raydir = torch.matmul(cam_pose[:, None, None, :3, :3], dirs.unsqueeze(-1))[0, :, :, :, 0]
        viewdirs = cam_raydir / cam_raydir.norm(dim=-1, keepdim=True)
        nears = torch.tensor(0.01).view(1, 1, 1).expand(H, W, -1)
        fars = sys.modules["nerfmatch.nerf.scene_utils"].rays_intersect_sphere(cam_centers.reshape(-1, 3), viewdirs.reshape(-1, 3), r=1).view(H, W, -1)
        dx = torch.sqrt(torch.sum((viewdirs[:-1] - viewdirs[1:]) ** 2, -1))
        dx = torch.cat([dx, dx[-2:-1]], 0)
        radii = dx[..., None] * 2 / np.sqrt(12)
        rays = torch.cat((cam_centers, viewdirs, nears, fars, viewdirs, radii), dim=-1)
        rays = rays[ds // 2 :: ds, ds // 2 :: ds].reshape((-1, rays.shape[-1]))
        n = rays.shape[0]
        g = torch.Generator().manual_seed(seed)
        img_ds = torch.rand(n, 3, generator=g)
        # the two passes (nerfmatch_evaluator.py:345-420)
        rays_d, viewd = rays[..., 3:6], rays[..., 8:11]
        z_vals, weights = None, None
        torch.manual_seed(7)
        keep = {}
        for key, model in (("coarse", ren.nerf_coarse), ("fine", ren.nerf_fine)):
            pts, z_vals = r.sample_smth_along_rays(rays.detach(), num_pts=128, z_vals=z_vals, weights=weights, use_disp=False, perturb=False,
                                                   embed_type="mip", model_type=key, scale_var=1)
            vflat = viewd[..., None, :].expand(*pts[0].shape[:-1], 3)
            mu = (z_vals[..., :-1] + z_vals[..., 1:]) / 2
            hw = (z_vals[..., :-1] - z_vals[..., 1:]) / 2
            eps = torch.tensor(torch.finfo(torch.float32).eps)
            t_mean = mu + (2 * mu * hw**2) / torch.maximum(eps, 3 * mu**2 + hw**2)
            var_flat = pts[1].contiguous().view(-1, 3)
            pts_z = rays[:, None, :3].expand(*pts[0].shape[:-1], 3) + t_mean[:, :, None] * vflat
            inputs = ren.xyz_encoder(pts_z.contiguous().view(-1, 3), y=var_flat)[0]
            inputs = torch.cat((inputs, ren.dirs_encoder(vflat.contiguous().view(-1, 3))), dim=-1)
            if key == "coarse":
                with torch.no_grad():
                    raw_outs, feats = model.forward(inputs, ret_pfeat=1, val=True)
            else:
                inputs.retain_grad()
                raw_outs, feats = model.forward(inputs, ret_pfeat=1, val=True)
                raw_outs.retain_grad()
                keep.update(inputs=inputs, raw=raw_outs, z=z_vals.detach().clone(), var=pts[1].detach().clone())
            raw_outs = raw_outs.view(-1, 128, raw_outs.shape[-1])
            rgb_map, _, acc_map, weights, depth_map, _ = r.volume_render_radiance_field(
                raw_outs, z_vals, rays_d, noise_std=0.0, white_bg=True, embed_type="mip", input_dim=4)
        loss = torch.nn.MSELoss(reduction="mean")(rgb_map, img_ds)
        loss.backward()
        out.update({
            f"{tag}.seed_gain_bias": np.array([seed, gain, bias], np.float64), f"{tag}.pose": cam_pose.detach()[0].numpy(),
            f"{tag}.hwds": np.array([H, W, ds]), f"{tag}.K": K.numpy(), f"{tag}.img": img_ds.numpy(),
            f"{tag}.z_fine": keep["z"].numpy(), f"{tag}.var_fine": keep["var"].numpy(), f"{tag}.rays": rays.detach().numpy(),
            f"{tag}.rgb_map": rgb_map.detach().numpy(), f"{tag}.loss": np.array(loss.item(), np.float64),
            f"{tag}.grad_pose": cam_pose.grad[0].numpy(), f"{tag}.grad_raw": keep["raw"].grad.numpy(),
            f"{tag}.grad_inputs_head": keep["inputs"].grad.numpy()[:1024],     # first 8 rays' samples
        })
        print(tag, "loss", loss.item(), "|grad|", cam_pose.grad.abs().max().item(), "acc", acc_map.mean().item())
    torch.set_grad_enabled(False)
    return out


def _stub_heavy_imports():
    import importlib

    class _Any(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            m = _Any(self.__name__ + "." + name)
            setattr(self, name, m)
            return m

        def __call__(self, *a, **k):
            return self

    for m in ("pytorch_lightning", "timm", "kornia", "transforms3d", "transforms3d.quaternions"):
        if m not in sys.modules:
            try:
                importlib.import_module(m)
            except Exception:
                sys.modules[m] = _Any(m)


def cache_cases(r):
    """Scene-cache file format, both directions:
      * a file written the way the reference writes it (nerf_evaluator.py:358-372: ``np.save(scene_dir / f"{img_idx}.npy",
        scene_pts)``), committed as tests/golden/ref_cache/<img_idx>.npy, which the repo's reader must load;
      * a file written by the repo's writer (nerfmatch_b200.cache.write_frame) read back with the reference's own
        ``load_frame_3d`` (datasets/data_loading.py:36-80); what the reference returned is stored in the fixture."""
    import importlib
    import tempfile
    from pathlib import Path

    import torch

    from nerfmatch_b200.cache import write_frame
    from oracle.weights import make_unnorm_scene, uniform01

    _stub_heavy_imports()
    pkg = types.ModuleType("nerfmatch.datasets")            # the package __init__ pulls in the training datasets: load the one module only
    pkg.__path__ = [os.path.join(REF, "nerfmatch", "datasets")]
    sys.modules.setdefault("nerfmatch.datasets", pkg)
    dl = importlib.import_module("nerfmatch.datasets.data_loading")

    n = 24
    pt3d = (uniform01(9, 1, n * 3).reshape(n, 3) * 4 - 2).astype(np.float32)
    feat = uniform01(9, 2, n * 256).reshape(n, 256).astype(np.float32)
    color = (uniform01(9, 3, n * 3).reshape(n, 3) * 1.2 - 0.1).astype(np.float32)
    un = make_unnorm_scene()
    out = {"pt3d": pt3d, "pt_feat": feat, "pt_color_raw": color, "unnorm_scene": un}
    frame = {"file_path": "seq-01/frame-000007.color.png"}
    img_idx = "seq-01_frame-000007"
    # (1) reference-style writer -> committed file
    ref_dir = Path(os.path.dirname(HERE)) / "tests" / "golden" / "ref_cache"
    ref_dir.mkdir(parents=True, exist_ok=True)
    scene_pts = dict(pt3d=torch.from_numpy(pt3d).numpy(), unnorm_scene=torch.from_numpy(un).numpy(), pt_feat=torch.from_numpy(feat).cpu().numpy(),
                     pt_color=torch.from_numpy(color).reshape(-1, 3).clamp(0, 1).cpu().numpy())
    np.save(ref_dir / f"{img_idx}.npy", scene_pts)
    a = dl.load_frame_3d(frame, str(ref_dir))
    assert np.array_equal(a[0], pt3d) and np.array_equal(a[1], feat) and a[2].all()
    # (2) repo writer -> reference reader
    with tempfile.TemporaryDirectory() as td:
        write_frame(td, img_idx, pt3d, un, feat, color)
        b = dl.load_frame_3d(frame, td, return_pose=True)
        item = np.load(os.path.join(td, f"{img_idx}.npy"), allow_pickle=True).item()
    out["ref_reads_ours.pt3d"], out["ref_reads_ours.pt_feat"], out["ref_reads_ours.mask"], out["ref_reads_ours.unnorm_scene"] = b[0], b[1], b[2], b[3]
    out["ref_reads_ours.c2w_is_none"] = np.array(b[4] is None)
    out["ref_reads_ours.keys"] = np.array(sorted(item))
    out["ref_reads_ours.pt_color"] = item["pt_color"]
    out["ref_written.keys"] = np.array(sorted(scene_pts))
    return out


def train_cases(r):
    """Training-mode forward: the reference's NerfRenderer.forward(validation=False) with
    set_training_mode(True) - perturbed sampling, sigma noise, s_fine / weights_fine outputs (renderer.py:217, 242, 283-286;
    nerf_trainer.py:140-147) - on CPU, with the random draws it consumed replayed in the same order from the same seed."""
    import torch

    from oracle.weights import K_7SCENES, make_pose, make_state_dict

    torch.set_grad_enabled(False)
    out = {}
    for tag, embed, pfeat, seed in (("mip", "mip", False, 31), ("normal", "normal", True, 32)):
        c = dict(embed=embed, nc=64, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False, sigma_target=(2.0, 1.0), ds=60, seed=seed)
        rays0 = r.sample_nerf_rays(480, 640, torch.from_numpy(K_7SCENES), torch.from_numpy(make_pose(seed)), ds=c["ds"], embed_type=embed)
        c["sigma_gain"], c["sigma_bias"] = calibrate_sigma(c, rays0.numpy())
        torch.manual_seed(0)
        ren = r.NerfRenderer(ref_config(embed, 64, 128, False, False), num_frames=None, training=True, stop_layer=3)
        sd = make_state_dict(seed, embed_type=embed, sigma_gain=c["sigma_gain"], sigma_bias=c["sigma_bias"])
        ren.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        ren.ret_pfeat = pfeat
        ren.set_training_mode(True)
        rays = rays0
        n = rays.shape[0]
        torch.manual_seed(13)
        preds = ren.forward(rays, step=0, ray_id=None, validation=False)
        assert "pts_fine" not in preds and "s_fine" in preds and "weights_fine" in preds
        # replay the draws in the order the reference consumed them (all on the CPU generator here)
        torch.manual_seed(13)
        if embed == "mip":
            t_rand = torch.rand(n, 65)
            noise_c = torch.randn(n, 64) * 1.0
            u = torch.empty(n, 65).uniform_(to=(1 / 65 - torch.finfo(torch.float32).eps))
            noise_f = torch.randn(n, 64) * 1.0
        else:
            t_rand = torch.rand(n, 64)
            noise_c = torch.randn(n, 64) * 1.0
            u = torch.rand(n, 128)
            noise_f = torch.randn(n, 192) * 1.0
        out.update({f"{tag}.rays": rays.numpy(), f"{tag}.t_rand": t_rand.numpy(), f"{tag}.u": u.numpy(), f"{tag}.noise_coarse": noise_c.numpy(),
                    f"{tag}.noise_fine": noise_f.numpy(), f"{tag}.case": np.array(repr(c)), f"{tag}.keys": np.array(list(preds.keys()))})
        for k, v in preds.items():
            out[f"{tag}.{k}"] = v.numpy()
        print(tag, "keys", list(preds.keys()), "acc_f", preds["weights_fine"].sum(-1).mean().item())
    return out


def main():
    r = import_reference()
    gdir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only:
            continue
        out = run_case(r, name, c)
        path = os.path.join(gdir, f"render_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "n =", out["rays"].shape[0], "->", path, os.path.getsize(path) // 1024, "KiB",
              "acc_c %.3f acc_f %.3f" % (out["weights_coarse"].sum(-1).mean(), out["weights_fine"].sum(-1).mean()),
              "wmax_f %.3f" % out["weights_fine"].max(-1).mean(), str(out["case"]))
    for fname, fn in (("raygen.npz", raygen_cases), ("stages.npz", stage_cases), ("match.npz", lambda _r: match_cases()),
                      ("rayhelpers.npz", rayhelper_cases), ("inerf_grad.npz", inerf_cases), ("cache_format.npz", cache_cases), ("train_forward.npz", train_cases)):
        if only and fname not in only:
            continue
        out = fn(r)
        path = os.path.join(gdir, fname)
        np.savez_compressed(path, **out)
        print(fname, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
