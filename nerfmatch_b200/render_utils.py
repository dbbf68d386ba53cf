"""Free functions other reference modules import from ``nerfmatch.nerf.render_utils`` /
``scene_utils``, re-created on top of the CUDA stage kernels.  Same names,
argument meaning and return layout; tensors must live on a CUDA device."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .autograd import CompositeFn, wants_grad
from .engine import Engine, EngineConfig

_engines = {}


def _engine(device, **kw) -> Engine:
    cfg = EngineConfig(**kw)
    key = (str(torch.device(device)), tuple(sorted(vars(cfg).items())))
    if key not in _engines:
        _engines[key] = Engine(cfg, device)
    return _engines[key]


def _need_cuda(t, what):
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise RuntimeError(f"{what}: nerfmatch_b200 works on CUDA tensors only (no CPU fallback)")


def sample_nerf_rays(H, W, K, c2w, ds=8, embed_type="mip", device=None):
    """render_utils.py:56-78.  K [3,3] and c2w [4,4] may be host tensors / arrays: only the pose is
    uploaded and the rays [n, 12|11] are generated on ``device`` (default: current CUDA device)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    eng = _engine(device, embed_type=embed_type)
    rays, _, missed = eng.raygen(torch.as_tensor(K).cpu(), torch.as_tensor(c2w).cpu(), H, W, ds, check_far=True)
    if missed:
        print("Fail to find far plane: camera outside the unit sphere! Set far to 1.")
    return rays


def get_K(H, W, focal_x, focal_y):
    """render_utils.py:15-20 (a 3x3 host tensor, as in the reference)."""
    return torch.tensor([[focal_x, 0, 0.5 * W], [0, focal_y, 0.5 * H], [0, 0, 1]], dtype=torch.float32)


def get_ray_dirs(H, W, K, flipped_yz=False, return_xys=False, device=None):
    """render_utils.py:23-34: camera-frame directions of all pixels, [H, W, 3] on ``device``."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    dirs = _engine(device).ray_dirs(torch.as_tensor(K).detach().cpu(), int(H), int(W), flipped_yz)
    if return_xys:
        ys, xs = torch.meshgrid(torch.arange(int(H), device=device), torch.arange(int(W), device=device), indexing="ij")
        return dirs, torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).float()
    return dirs


def get_rays_c2w(dirs, c2w):
    """render_utils.py:37-41: (rays_o, rays_d, viewdirs), each shaped like ``dirs``."""
    _need_cuda(dirs, "get_rays_c2w")
    o, d, v = _engine(dirs.device).rays_c2w(dirs, c2w)
    return o, d, v


def prepare_rays_data(rays_o, rays_d, viewdirs=None, near=0.0, far=1.0, flatten=True, comp_radii=False):
    """render_utils.py:81-104: rays = cat(o, d, near, far[, viewdirs][, radii]); ``rays_d`` is [H, W, 3] (the radii are
    finite differences along the image rows) or [n, 3] without radii."""
    _need_cuda(rays_d, "prepare_rays_data")
    if rays_d.dim() == 3:
        H, W = rays_d.shape[:2]
    else:
        if comp_radii:
            raise ValueError("comp_radii needs rays_d of shape [H, W, 3]")
        H, W = 1, rays_d.shape[0]
    rays = _engine(rays_d.device).prepare_rays(rays_o, rays_d, viewdirs, near, far, H, W, comp_radii)
    return rays if flatten else rays.reshape(*rays_d.shape[:-1], rays.shape[-1])


def prepare_rays_from_pose(c2w, K, near=0.0, far=1.0, flipped_yz=False, comp_radii=False, device=None):
    """render_utils.py:44-53: full-resolution rays of one pose (un-normalised rays_d, constant near / far)."""
    K = torch.as_tensor(K).detach().cpu()
    w, h = (K[:2, 2].numpy().astype(np.int32) * 2).tolist()
    dirs = get_ray_dirs(h, w, K, flipped_yz=flipped_yz, device=device)
    rays_o, rays_d, viewdirs = get_rays_c2w(dirs, torch.as_tensor(c2w))
    return prepare_rays_data(rays_o, rays_d, viewdirs, near, far, comp_radii=comp_radii).float()


def rays_intersect_sphere(rays_o, rays_d, r=1):
    """scene_utils.py:101-120 - tiny and only used as a host-side helper by the reference datasets;
    kept as a torch expression on whatever device the inputs live on."""
    odotd = torch.sum(rays_o * rays_d, 1)
    d2 = torch.sum(rays_d**2, 1)
    o2 = torch.sum(rays_o**2, 1)
    det = odotd**2 + (r**2 - o2) * d2
    assert torch.all(det >= 0), "Not all your cameras are bounded by the unit sphere; please make sure the cameras are normalized properly!"
    return (torch.sqrt(det) - odotd) / d2


def sample_smth_along_rays(rays, num_pts=64, z_vals=None, weights=None, use_disp=True, perturb=False,
                           embed_type="normal", model_type="coarse", randomized=True, resample_padding=0.01,
                           scale_var=-1, num_pts_coarse=None):
    """render_utils.py:267-323.  Returns ((mean, var), t_vals) in mip mode and (pts, z_vals) otherwise.
    Random draws are made with torch exactly as the reference does, then injected into the kernel."""
    _need_cuda(rays, "sample_smth_along_rays")
    if perturb and embed_type != "mip":
        raise NotImplementedError("perturb=True (training) is outside the built hot path")
    rays = rays.detach().float().contiguous()
    n = rays.shape[0]
    mip = embed_type == "mip"
    if model_type == "coarse" or z_vals is None:
        eng = _engine(rays.device, embed_type=embed_type, num_pts_coarse=num_pts, use_disp=bool(use_disp) and not mip,
                      mip_var_scale=float(scale_var), resample_padding=resample_padding)
        t_rand = torch.rand(n, num_pts + 1, device=rays.device) if (mip and randomized) else None
        z, pts, var = eng.stage_sample_coarse(rays, t_rand)
        return ((pts, var), z) if mip else (pts, z)
    sc = z_vals.shape[-1] - 1 if mip else z_vals.shape[-1]
    eng = _engine(rays.device, embed_type=embed_type, num_pts_coarse=sc, num_pts_fine=num_pts,
                  mip_var_scale=float(scale_var), resample_padding=resample_padding)
    u_jit = None
    if mip and randomized:
        S1 = sc + 1
        u_jit = torch.empty(n, S1, device=rays.device).uniform_(to=(1 / S1 - float(torch.finfo(torch.float32).eps)))
    z, idx, pts, var = eng.stage_resample(rays, z_vals.detach().float().contiguous(), weights.detach().float().contiguous(), u_jit)
    return ((pts, var), z) if mip else (pts, z)


def volume_render_radiance_field(radiance_field, z_vals, rays_d, noise_std=0.0, white_bg=True, embed_type=False,
                                 out_last=False, input_dim=4):
    """render_utils.py:176-230 -> (rgb_map, disp_map, acc_map, weights, depth_map, None)."""
    _need_cuda(radiance_field, "volume_render_radiance_field")
    if noise_std > 0.0 or out_last or input_dim != 4:
        raise NotImplementedError("sigma noise / out_last heads (training) are outside the built hot path")
    mip = str(embed_type)[:3] == "mip"
    eng = _engine(radiance_field.device, embed_type="mip" if mip else "normal")
    flags = L.NMR_FLAG_WHITE_BG if white_bg else 0
    if wants_grad(radiance_field, rays_d):      # iNeRF refinement (nerfmatch_evaluator.py:410-420); z_vals is a constant there
        rgb, disp, acc, weights, depth = CompositeFn.apply(radiance_field[..., :4], rays_d, z_vals, eng, flags)
        return rgb, disp, acc, weights, depth, None
    raw = radiance_field[..., :4].detach().float().contiguous()
    out = eng.stage_composite(raw, z_vals.detach().float().contiguous(), rays_d.detach().float().contiguous(), flags=flags)
    return out["rgb"], out["disp"], out["acc"], out["weights"], out["depth"], None


def unnormaliz_pts(pt3d_normed, unnorm_mat):
    """utils/geometry.py:76-85 for callers that hold points already; the renderer folds this affine into
    the compositing kernel's epilogue instead."""
    R, t = unnorm_mat[..., :3, :3], unnorm_mat[..., :3, 3]
    return pt3d_normed @ R.transpose(-1, -2) + t[..., None, :]
