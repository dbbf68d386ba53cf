"""Thin owner of one ``nmr_handle``: creation from a configuration, weight upload, and typed wrappers
around the C-ABI calls that take torch tensors (device memory + stream plumbing only)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib as L


@dataclass
class EngineConfig:
    embed_type: str = "mip"
    xyz_num_freqs: int = 15
    dirs_num_freqs: int = 4
    num_pts_coarse: int = 64
    num_pts_fine: int = 128
    hid_dim: int = 256
    layer_num: int = 8
    skips: tuple = (4,)
    app_dim: int = 0
    stop_layer_coarse: int = -1
    stop_layer_fine: int = -1
    use_disp: bool = False
    precision: str = "fp32"
    mip_var_scale: float = -1.0
    resample_padding: float = 0.01
    max_chunk_points: int = 0

    @property
    def mip(self) -> bool:
        return self.embed_type == "mip"

    @property
    def xyz_dim(self) -> int:
        return 6 * self.xyz_num_freqs + (0 if self.mip else 3)

    @property
    def dirs_dim(self) -> int:
        return 6 * self.dirs_num_freqs + 3

    @property
    def in_dim(self) -> int:
        return self.xyz_dim + self.dirs_dim + self.app_dim

    @property
    def s_fine(self) -> int:
        return self.num_pts_coarse if self.mip else self.num_pts_coarse + self.num_pts_fine

    @property
    def z_coarse(self) -> int:
        return self.num_pts_coarse + 1 if self.mip else self.num_pts_coarse

    @property
    def z_fine(self) -> int:
        return self.num_pts_coarse + 1 if self.mip else self.s_fine

    @property
    def idx_cols(self) -> int:
        return self.num_pts_coarse + 1 if self.mip else self.num_pts_fine


def _ptr(t):
    """device pointer of a tensor, or a raw address (int: e.g. a peer-memory pointer from ``Engine.peer_open``)"""
    if t is None:
        return None
    return C.c_void_p(int(t)) if isinstance(t, int) else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


_NET_KEYS = [f"pts_linears.{i}" for i in range(8)] + ["views_linears.0", "feature_linear", "alpha_linear", "rgb_linear"]


class Engine:
    """One nmr_handle bound to one CUDA device."""

    def __init__(self, cfg: EngineConfig, device: torch.device):
        if len(tuple(cfg.skips)) != 1:
            raise L.NmrError(f"only skips=[4] is built (got {cfg.skips})")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.NmrError("the render path runs on CUDA only (no CPU fallback)")
        self.lib = L.load()
        c = L.NmrConfig(
            abi_version=L.ABI_VERSION,
            embed_type=L.NMR_EMBED_MIP if cfg.mip else L.NMR_EMBED_NORMAL,
            xyz_num_freqs=cfg.xyz_num_freqs, dirs_num_freqs=cfg.dirs_num_freqs,
            num_pts_coarse=cfg.num_pts_coarse, num_pts_fine=cfg.num_pts_fine,
            hid_dim=cfg.hid_dim, layer_num=cfg.layer_num, skip_layer=int(tuple(cfg.skips)[0]),
            app_dim=cfg.app_dim, stop_layer_coarse=cfg.stop_layer_coarse, stop_layer_fine=cfg.stop_layer_fine,
            use_disp=int(cfg.use_disp), precision=self._prec(cfg.precision),
            mip_var_scale=float(cfg.mip_var_scale), resample_padding=float(cfg.resample_padding),
            max_chunk_points=int(cfg.max_chunk_points),
        )
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            st = self.lib.nmr_create(C.byref(c), C.byref(h))
        L.check(None, st, "nmr_create")
        self.h = h
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.nmr_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- plumbing -------------------------------------------------------------------------
    @staticmethod
    def _prec(name: str) -> int:
        if name not in L.PRECISIONS:
            raise L.NmrError(f"precision must be one of {sorted(L.PRECISIONS)}, got {name!r}")
        return L.PRECISIONS[name]

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def check(self, st, what):
        L.check(self.h, st, what)

    def set_precision(self, precision: str):
        self.check(self.lib.nmr_set_precision(self.h, self._prec(precision)), "nmr_set_precision")
        self.cfg.precision = precision

    def set_weights(self, which: int, sd: Dict[str, torch.Tensor]):
        """sd: reference state_dict of ONE net (keys without the nerf_coarse./nerf_fine. prefix)."""
        keep = {}
        w = L.NmrWeightPtrs()
        for k in _NET_KEYS:
            for part in ("weight", "bias"):
                keep[(k, part)] = _f32c(sd[f"{k}.{part}"], self.device)
        for i in range(8):
            w.pts_w[i] = keep[(f"pts_linears.{i}", "weight")].data_ptr()
            w.pts_b[i] = keep[(f"pts_linears.{i}", "bias")].data_ptr()
        w.views_w, w.views_b = keep[("views_linears.0", "weight")].data_ptr(), keep[("views_linears.0", "bias")].data_ptr()
        w.feature_w, w.feature_b = keep[("feature_linear", "weight")].data_ptr(), keep[("feature_linear", "bias")].data_ptr()
        w.alpha_w, w.alpha_b = keep[("alpha_linear", "weight")].data_ptr(), keep[("alpha_linear", "bias")].data_ptr()
        w.rgb_w, w.rgb_b = keep[("rgb_linear", "weight")].data_ptr(), keep[("rgb_linear", "bias")].data_ptr()
        exp_in = {0: self.cfg.xyz_dim, 5: self.cfg.xyz_dim + 256}
        for i in range(8):
            shp = tuple(keep[(f"pts_linears.{i}", "weight")].shape)
            if shp != (256, exp_in.get(i, 256)):
                raise L.NmrError(f"pts_linears.{i}.weight has shape {shp}, expected {(256, exp_in.get(i, 256))}")
        vshape = tuple(keep[("views_linears.0", "weight")].shape)
        if vshape != (128, 256 + self.cfg.dirs_dim + self.cfg.app_dim):
            raise L.NmrError(f"views_linears.0.weight has shape {vshape}")
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_set_weights(self.h, which, C.byref(w))
        self.check(st, "nmr_set_weights")

    def set_appearance(self, table: torch.Tensor):
        t = _f32c(table, self.device)
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_set_appearance(self.h, _ptr(t), t.shape[0])
        self.check(st, "nmr_set_appearance")

    def launch_count(self) -> int:
        return int(self.lib.nmr_launch_count(self.h))

    HOST_OUTPUTS = ("feat_coarse", "feat_fine", "pts_coarse", "pts_fine", "rgb_coarse", "rgb_fine", "depth_coarse", "depth_fine", "pt3d")

    def copy_host_staging(self, name: str, dst, nbytes: Optional[int] = None):
        """Device copy of output ``name`` of the last render_host call -> ``dst`` (device tensor, or a raw / peer address with
        ``nbytes``), device to device."""
        nbytes = dst.numel() * dst.element_size() if nbytes is None else nbytes
        with torch.cuda.device(self.device):
            st = self.lib.nmr_copy_host_staging(self.h, self.HOST_OUTPUTS.index(name), _ptr(dst), nbytes, self.stream())
        self.check(st, "nmr_copy_host_staging")
        return dst

    def set_fuse_composite(self, on: bool):
        """bf16 path: tile-local feature compositing inside the MLP kernel (default on) vs. per-sample features through HBM."""
        self.check(self.lib.nmr_set_fuse_composite(self.h, int(on)), "nmr_set_fuse_composite")

    def set_timing(self, on: bool):
        self.check(self.lib.nmr_set_timing(self.h, int(on)), "nmr_set_timing")

    def last_stage_ms(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self.check(self.lib.nmr_last_stage_ms(self.h, C.byref(a), C.byref(b), C.byref(c)), "nmr_last_stage_ms")
        return {"sample": a.value, "mlp": b.value, "composite": c.value}

    # -- hot path -------------------------------------------------------------------------
    def raygen(self, K, c2w, H: int, W: int, ds: int, check_far: bool = False):
        K = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(K, dtype=torch.float32).reshape(-1).tolist()])
        c2w_l = torch.as_tensor(c2w, dtype=torch.float32)[:3, :4].reshape(-1).tolist()
        c2w_c = (C.c_float * 12)(*[float(v) for v in c2w_l])
        h, w = len(range(ds // 2, H, ds)), len(range(ds // 2, W, ds))
        rays = torch.empty(h * w, 12 if self.cfg.mip else 11, device=self.device, dtype=torch.float32)
        flag = C.c_int32(0)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_raygen(self.h, K, c2w_c, H, W, ds, _ptr(rays), C.byref(flag) if check_far else None, self.stream())
        self.check(st, "nmr_raygen")
        return rays, (h, w), bool(flag.value)

    # ---- the reference's ray helper functions as device ops (render_utils.py:23-41, 81-104) ----
    def ray_dirs(self, K, H: int, W: int, flipped_yz: bool = False):
        Kc = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(K, dtype=torch.float32).reshape(-1).tolist()])
        dirs = torch.empty(H, W, 3, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_ray_dirs(self.h, Kc, H, W, int(bool(flipped_yz)), _ptr(dirs), self.stream())
        self.check(st, "nmr_ray_dirs")
        return dirs

    def rays_c2w(self, dirs, c2w):
        c2w_c = (C.c_float * 12)(*[float(v) for v in torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()[:3, :4].reshape(-1).tolist()])
        dirs = dirs.detach().float().contiguous()
        o, d, v = torch.empty_like(dirs), torch.empty_like(dirs), torch.empty_like(dirs)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_rays_c2w(self.h, _ptr(dirs), dirs.numel() // 3, c2w_c, _ptr(o), _ptr(d), _ptr(v), self.stream())
        self.check(st, "nmr_rays_c2w")
        return o, d, v

    def prepare_rays(self, rays_o, rays_d, viewdirs, near, far, H: int, W: int, comp_radii: bool):
        f = lambda t: None if t is None else t.detach().float().contiguous()
        rays_o, rays_d, viewdirs = f(rays_o), f(rays_d), f(viewdirs)
        near_t = f(near) if isinstance(near, torch.Tensor) else None
        far_t = f(far) if isinstance(far, torch.Tensor) else None
        ncols = 8 + (3 if viewdirs is not None else 0) + (1 if comp_radii else 0)
        out = torch.empty(H * W, ncols, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_prepare_rays(self.h, _ptr(rays_o), _ptr(rays_d), _ptr(viewdirs), _ptr(near_t),
                                           0.0 if near_t is not None else float(near), _ptr(far_t),
                                           0.0 if far_t is not None else float(far), H, W, int(bool(comp_radii)), _ptr(out), self.stream())
        self.check(st, "nmr_prepare_rays")
        return out

    def make_args(self, n: int, ncols: int, rays=None, ray_id=None, t_rand=None, u_jitter=None, flags=0,
                  rng_mode=L.NMR_RNG_NONE, seed=0, unnorm=None, outs: Optional[Dict[str, torch.Tensor]] = None,
                  sigma_noise_coarse=None, sigma_noise_fine=None):
        a = L.NmrRenderArgs()
        a.rays, a.n, a.ncols, a.flags = _ptr(rays), n, ncols, flags
        a.ray_id, a.t_rand, a.u_jitter = _ptr(ray_id), _ptr(t_rand), _ptr(u_jitter)
        a.sigma_noise_coarse, a.sigma_noise_fine = _ptr(sigma_noise_coarse), _ptr(sigma_noise_fine)
        a.rng_mode, a.seed = rng_mode, seed
        keep = [rays, ray_id, t_rand, u_jitter, sigma_noise_coarse, sigma_noise_fine]
        if unnorm is not None:
            un = (C.c_float * 12)(*[float(v) for v in torch.as_tensor(unnorm, dtype=torch.float32)[:3, :4].reshape(-1).tolist()])
            a.unnorm = C.cast(un, C.c_void_p)
            keep.append(un)
        for k, t in (outs or {}).items():       # tensors, or raw (peer) addresses
            setattr(a, k, _ptr(t))
            keep.append(t)
        return a, keep

    def alloc_outputs(self, n: int, want_feat: bool, want_pt3d: bool, intermediates: bool = False, device=None, pin=False,
                      training: bool = False):
        cfg = self.cfg
        dev = self.device if device is None else device

        def e(*shape, dtype=torch.float32):
            t = torch.empty(*shape, dtype=dtype, device=dev)
            return t.pin_memory() if pin else t

        o = {}
        for key in ("coarse", "fine"):
            if want_feat:
                o[f"feat_{key}"] = e(n, 256)
            o[f"pts_{key}"] = e(n, 3)
            o[f"rgb_{key}"] = e(n, 3)
            o[f"depth_{key}"] = e(n)
        if want_pt3d:
            o["pt3d"] = e(n, 3)
        if training:         # forward(validation=False): s_fine / weights_fine (renderer.py:283-286)
            o["s_fine"], o["weights_fine"] = e(n, cfg.z_fine), e(n, cfg.s_fine)
        if intermediates:
            o["z_coarse"], o["z_fine"] = e(n, cfg.z_coarse), e(n, cfg.z_fine)
            o["weights_coarse"], o["weights_fine"] = e(n, cfg.num_pts_coarse), e(n, cfg.s_fine)
            o["raw_coarse"], o["raw_fine"] = e(n, cfg.num_pts_coarse, 4), e(n, cfg.s_fine, 4)
            o["idx_fine"] = e(n, cfg.idx_cols, dtype=torch.int32)
        return o

    def render(self, args: L.NmrRenderArgs):
        with torch.cuda.device(self.device):
            st = self.lib.nmr_render(self.h, C.byref(args), self.stream())
        self.check(st, "nmr_render")

    def render_host(self, args: L.NmrRenderArgs):
        with torch.cuda.device(self.device):
            st = self.lib.nmr_render_host(self.h, C.byref(args), self.stream())
        self.check(st, "nmr_render_host")

    def render_view(self, K, c2w, H, W, ds, args: L.NmrRenderArgs):
        Kc = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(K, dtype=torch.float32).reshape(-1).tolist()])
        cc = (C.c_float * 12)(*[float(v) for v in torch.as_tensor(c2w, dtype=torch.float32)[:3, :4].reshape(-1).tolist()])
        with torch.cuda.device(self.device):
            st = self.lib.nmr_render_view(self.h, Kc, cc, H, W, ds, C.byref(args), self.stream())
        self.check(st, "nmr_render_view")

    # -- stages ---------------------------------------------------------------------------
    def stage_sample_coarse(self, rays, t_rand=None, want_pts=True):
        cfg = self.cfg
        n = rays.shape[0]
        z = torch.empty(n, cfg.z_coarse, device=self.device)
        pts = torch.empty(n, cfg.num_pts_coarse, 3, device=self.device) if want_pts else None
        var = torch.empty(n, cfg.num_pts_coarse, 3, device=self.device) if (want_pts and cfg.mip) else None
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_sample_coarse(self.h, _ptr(rays), n, rays.shape[1], _ptr(t_rand), _ptr(z), _ptr(pts), _ptr(var), self.stream())
        self.check(st, "nmr_stage_sample_coarse")
        return z, pts, var

    def stage_resample(self, rays, z_in, weights, u_jitter=None, want_pts=True):
        cfg = self.cfg
        n = rays.shape[0]
        z = torch.empty(n, cfg.z_fine, device=self.device)
        idx = torch.empty(n, cfg.idx_cols, device=self.device, dtype=torch.int32)
        pts = torch.empty(n, cfg.s_fine, 3, device=self.device) if want_pts else None
        var = torch.empty(n, cfg.s_fine, 3, device=self.device) if (want_pts and cfg.mip) else None
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_resample(self.h, _ptr(rays), n, rays.shape[1], _ptr(z_in), _ptr(weights), _ptr(u_jitter),
                                             _ptr(z), _ptr(idx), _ptr(pts), _ptr(var), self.stream())
        self.check(st, "nmr_stage_resample")
        return z, idx, pts, var

    def stage_encode(self, kind: int, x, var=None):
        cfg = self.cfg
        P = x.shape[0]
        F = cfg.xyz_num_freqs if kind == 0 else cfg.dirs_num_freqs
        dim = 6 * F if (cfg.mip and kind == 0 and var is not None) else 6 * F + 3
        out = torch.empty(P, dim, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_encode(self.h, kind, _ptr(x), _ptr(var), P, _ptr(out), self.stream())
        self.check(st, "nmr_stage_encode")
        return out

    def stage_mlp(self, which: int, x, stop_layer: int, want_feat=True):
        P = x.shape[0]
        raw = torch.empty(P, 4, device=self.device)
        feat = torch.empty(P, 256, device=self.device) if want_feat else None
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_mlp(self.h, which, _ptr(x), P, stop_layer, _ptr(raw), _ptr(feat), self.stream())
        self.check(st, "nmr_stage_mlp")
        return raw, feat

    def stage_mlp_probe(self, which: int, x, step: int):
        P = x.shape[0]
        out = torch.zeros(P, 256, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_mlp_probe(self.h, which, _ptr(x), P, step, _ptr(out), self.stream())
        self.check(st, "nmr_stage_mlp_probe")
        return out

    def stage_composite(self, raw, z, rays_d, flags=0, feats=None, pts=None):
        n, S = raw.shape[0], raw.shape[1]
        dev = self.device
        rgb, depth, acc, disp = torch.empty(n, 3, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev)
        weights = torch.empty(n, S, device=dev)
        feat_out = torch.empty(n, 256, device=dev) if feats is not None else None
        pts_out = torch.empty(n, 3, device=dev) if pts is not None else None
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_composite(self.h, _ptr(raw), _ptr(z), _ptr(rays_d), n, S, flags, _ptr(feats), _ptr(pts),
                                              _ptr(rgb), _ptr(depth), _ptr(acc), _ptr(disp), _ptr(weights), _ptr(feat_out),
                                              _ptr(pts_out), self.stream())
        self.check(st, "nmr_stage_composite")
        return dict(rgb=rgb, depth=depth, acc=acc, disp=disp, weights=weights, feat=feat_out, pts=pts_out)

    # -- backward stages (iNeRF refinement) ----------------------------
    def stage_encode_backward(self, kind: int, x, var, grad_out):
        P = x.shape[0]
        gx = torch.empty(P, 3, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_encode_backward(self.h, kind, _ptr(x), _ptr(var), P, _ptr(grad_out), _ptr(gx), self.stream())
        self.check(st, "nmr_stage_encode_backward")
        return gx

    def stage_mlp_backward(self, which: int, x, stop_layer: int, grad_raw, grad_feat=None, raw_fwd=None, masks=None):
        """x [P,in_dim] (recompute form) or raw_fwd + masks from ``stage_mlp_masks`` (x may be None)."""
        P = grad_raw.shape[0]
        gx = torch.empty(P, self.cfg.in_dim, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_mlp_backward(self.h, which, _ptr(x), P, stop_layer, _ptr(grad_raw), _ptr(grad_feat), _ptr(raw_fwd),
                                                 _ptr(masks), _ptr(gx), self.stream())
        self.check(st, "nmr_stage_mlp_backward")
        return gx

    def stage_mlp_masks(self, which: int, x, stop_layer: int):
        """Forward on the tcgen05 path that also returns the ReLU masks for ``stage_mlp_backward`` (int32 [tiles, 9216])."""
        P = x.shape[0]
        raw = torch.empty(P, 4, device=self.device)
        feat = torch.empty(P, 256, device=self.device)
        masks = torch.empty((P + 127) // 128, 9 * 8 * 128, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_mlp_masks(self.h, which, _ptr(x), P, stop_layer, _ptr(raw), _ptr(feat), _ptr(masks), self.stream())
        self.check(st, "nmr_stage_mlp_masks")
        return raw, feat, masks

    def stage_composite_backward(self, raw, z, rays_d, flags=0, grad_rgb=None, grad_depth=None, grad_acc=None, grad_weights=None,
                                 want_rays_d=True):
        n, S = raw.shape[0], raw.shape[1]
        g_raw = torch.empty(n, S, 4, device=self.device)
        g_d = torch.empty(n, 3, device=self.device) if want_rays_d else None
        with torch.cuda.device(self.device):
            st = self.lib.nmr_stage_composite_backward(self.h, _ptr(raw), _ptr(z), _ptr(rays_d), n, S, flags, _ptr(grad_rgb), _ptr(grad_depth),
                                                       _ptr(grad_acc), _ptr(grad_weights), _ptr(g_raw), _ptr(g_d), self.stream())
        self.check(st, "nmr_stage_composite_backward")
        return g_raw, g_d

    # -- peer-memory exchange (multi-GPU: outputs delivered into the matcher rank's memory; include/nmr.h) --------
    def peer_alloc(self, nbytes: int):
        ptr, hd = C.c_void_p(), (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            st = self.lib.nmr_peer_alloc(self.h, nbytes, C.byref(ptr), C.cast(hd, C.c_void_p))
        self.check(st, "nmr_peer_alloc")
        return int(ptr.value), bytes(hd)

    def peer_open(self, handle: bytes) -> int:
        ptr, hd = C.c_void_p(), (C.c_ubyte * 64)(*handle)
        with torch.cuda.device(self.device):
            st = self.lib.nmr_peer_open(self.h, C.cast(hd, C.c_void_p), C.byref(ptr))
        self.check(st, "nmr_peer_open")
        return int(ptr.value)

    def peer_close(self, ptr: int):
        self.check(self.lib.nmr_peer_close(self.h, C.c_void_p(ptr)), "nmr_peer_close")

    def peer_free(self, ptr: int):
        self.check(self.lib.nmr_peer_free(self.h, C.c_void_p(ptr)), "nmr_peer_free")

    def peer_signal(self, flag_ptr: int, value: int):
        with torch.cuda.device(self.device):
            st = self.lib.nmr_peer_signal(self.h, C.c_void_p(flag_ptr), value & 0xFFFFFFFF, self.stream())
        self.check(st, "nmr_peer_signal")

    def peer_wait(self, flags_ptr: int, count: int, value: int):
        with torch.cuda.device(self.device):
            st = self.lib.nmr_peer_wait(self.h, C.c_void_p(flags_ptr), count, value & 0xFFFFFFFF, self.stream())
        self.check(st, "nmr_peer_wait")

    def peer_wait_status(self) -> int:
        return int(self.lib.nmr_peer_wait_status(self.h))
