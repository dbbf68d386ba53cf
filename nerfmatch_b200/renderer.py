"""Drop-in replacement for ``nerfmatch.nerf.renderer.NerfRenderer`` (reference renderer.py:26-333).

Same constructor, attributes, ``state_dict`` keys and return dictionaries; the work is done by
libnmr.so (hand-written sm_100a CUDA) through the C ABI in include/nmr.h.  PyTorch is used for
device memory, the current stream, random draws and ``torch.distributed`` only.

Built: the inference path the evaluators use (``predict`` / ``forward(validation=True)`` /
``render_novel_view``) and the training-mode FORWARD (``forward(validation=False)``: perturbed
sampling, sigma noise, ``s_fine`` / ``weights_fine`` outputs; renderer.py:217, 242, 283-286) -
outputs carry no autograd graph w.r.t. the network weights (optimising the NeRF itself is outside the
hot-path scope).  ``out_scr`` heads, ``single_model`` and ``pfeat_mask`` raise ``NotImplementedError``.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .embedding import FourierEmbedding, PositionalEncodingMIP
from .engine import Engine, EngineConfig
from .models import NeRF

_EPS32 = float(torch.finfo(torch.float32).eps)


class NerfRenderer(nn.Module):
    def __init__(self, config, num_frames=None, training=True, stop_layer=-1, precision: Optional[str] = None,
                 max_chunk_points: int = 0):
        super().__init__()
        self.training = training
        r, e, dcfg = config.render, config.embedding, config.data
        self.chunksize = r.chunksize
        self.use_disp = r.use_disp
        self.perturb = r.perturb
        self.noise_std = r.noise_std
        self.white_bg = r.white_bg or getattr(dcfg, "white_bg", False)
        self.use_viewdirs = r.use_viewdirs
        self.embed_type = getattr(e, "type", "normal")
        self.bg_color = [1.0, 1.0, 1.0] if self.white_bg else None
        self.img_wh = dcfg.img_wh
        self.mip_var_scale = getattr(e, "mip_var_scale", -1)
        self.out_scr = getattr(dcfg, "out_scr", False)
        self.single_model = getattr(r, "single_model", False)
        self.appearance_embedding = getattr(e, "appearance_embed", False)
        if self.out_scr or not self.use_viewdirs:
            raise NotImplementedError("out_scr heads / use_viewdirs=False are outside the built hot path")
        if self.embed_type not in ("mip", "normal"):
            raise NotImplementedError(f"embedding.type={self.embed_type!r}")
        self.num_out_ch = 0

        mip = self.embed_type == "mip"
        enc = PositionalEncodingMIP if mip else FourierEmbedding
        self.xyz_encoder = enc(e.xyz_num_freqs)
        self.dirs_encoder = enc(e.dirs_num_freqs)
        xyz_dim = self.xyz_encoder.get_embedding_dim(3) - (3 if mip else 0)
        dirs_dim = self.dirs_encoder.get_embedding_dim(3)
        app_dim = 16 if self.appearance_embedding else 0

        def net_conf(base, stop):
            c = dict(vars(base))
            c.update(use_viewdirs=True, xyz_dim=xyz_dim, dirs_dim=dirs_dim, app_dim=app_dim, out_3d_pnt=False, out_add_ch=0)
            if stop is not None:
                c["stop_layer"] = stop
            return c

        if not self.single_model:          # renderer.py:76-87: with single_model the coarse net does not exist at all
            self.num_pts_coarse = config.coarse_nerf.num_pts
            self.nerf_coarse = NeRF(net_conf(config.coarse_nerf, None))   # stop_layer only reaches the fine net
        fine_conf = getattr(config, "fine_nerf", None)
        if fine_conf is None:
            raise NotImplementedError("a renderer without fine_nerf is outside the built hot path")
        self.num_pts_fine = fine_conf.num_pts
        self.nerf_fine = NeRF(net_conf(fine_conf, stop_layer))            # renderer.py:99
        self.output_dim = getattr(config.fine_nerf, "output_dim", 4)
        self.embedding_a = nn.Embedding(num_frames, 16) if self.appearance_embedding else None

        self.ret_pfeat = False
        self.pfeat_mask = None
        self.feat_comb = "lin"
        self.weight_dir = getattr(config.loss, "weight_dir", 1)

        # ---- B200 engine state (not part of the reference surface) ----
        self.precision = precision or os.environ.get("NMR_PRECISION", "bf16")
        self.rng = os.environ.get("NMR_RNG", "torch")   # "torch": same draws as the reference; "philox": in-kernel
        self.max_chunk_points = max_chunk_points
        self._engines = {}

    # ------------------------------------------------------------------------------------
    def set_training_mode(self, state):
        self.training = state

    def _nets(self):
        """(coarse-pass net, fine-pass net, coarse sample count): single_model runs the fine net in both passes with
        num_pts_fine samples each (renderer.py:191-201)."""
        if self.single_model:
            return self.nerf_fine, self.nerf_fine, self.num_pts_fine
        return self.nerf_coarse, self.nerf_fine, self.num_pts_coarse

    def _engine_config(self) -> EngineConfig:
        coarse, fine, sc = self._nets()
        return EngineConfig(
            embed_type=self.embed_type, xyz_num_freqs=self.xyz_encoder.num_freqs, dirs_num_freqs=self.dirs_encoder.num_freqs,
            num_pts_coarse=sc, num_pts_fine=self.num_pts_fine,
            app_dim=16 if self.appearance_embedding else 0,
            stop_layer_coarse=coarse.stop_layer, stop_layer_fine=fine.stop_layer,
            use_disp=bool(self.use_disp), precision=self.precision, mip_var_scale=float(self.mip_var_scale),
            max_chunk_points=self.max_chunk_points,
        )

    def engine(self, device) -> Engine:
        """The nmr_handle for ``device``; (re)uploads weights when a parameter changed
        (load_state_dict, .to(), in-place edits bump the tensors' version counters)."""
        key = str(torch.device(device))
        cfg = self._engine_config()
        # everything the nmr_handle fixes at creation (the reference reads these attributes live on every call)
        static = tuple(sorted((k, v) for k, v in vars(cfg).items() if k != "precision"))
        coarse, fine, _ = self._nets()
        weights = (coarse.weights_signature(), fine.weights_signature(),
                   None if self.embedding_a is None else (self.embedding_a.weight.data_ptr(), self.embedding_a.weight._version))
        ent = self._engines.get(key)
        if ent is None or ent[1] != (weights, static):
            eng = Engine(cfg, device) if (ent is None or ent[1][1] != static) else ent[0]
            eng.set_weights(L.NMR_NET_COARSE, coarse.net_state())
            eng.set_weights(L.NMR_NET_FINE, fine.net_state())
            if self.embedding_a is not None:
                eng.set_appearance(self.embedding_a.weight)
            ent = (eng, (weights, static))
            self._engines[key] = ent
        if ent[0].cfg.precision != self.precision:
            ent[0].set_precision(self.precision)
        return ent[0]

    def _flags(self):
        f = 0
        if self.ret_pfeat:
            f |= L.NMR_FLAG_RET_PFEAT
        if self.feat_comb == "max":
            f |= L.NMR_FLAG_FEAT_MAX
        if self.white_bg:
            f |= L.NMR_FLAG_WHITE_BG
        return f

    def _train_draws(self, n, device):
        """Training-mode randomness (``self.training``): the reference's draws, made with the same calls, shapes, devices and
        ORDER so that a seeded run consumes the generators identically (coarse sampling draw, coarse sigma noise, fine
        sampling draw, fine sigma noise):
          mip      torch.rand [n,Sc+1] (render_utils.py:444) . randn [n,Sc] (:191) . uniform_ [n,Sc+1] (:483-485) . randn [n,Sc]
          classic  torch.rand_like [n,Sc] (:122) . randn [n,Sc] . torch.rand [n,Nf] ON THE CPU (:247, then .to(device)) . randn [n,Sc+Nf]
        ``perturb`` gates the classic sampling draws only (mip jitters regardless); ``noise_std`` <= 0 skips the noise."""
        mip = self.embed_type == "mip"
        Sc, Nf = self._nets()[2], self.num_pts_fine
        noise = lambda S: (torch.randn(n, S, device=device) * self.noise_std) if self.noise_std > 0.0 else None
        if mip:
            t_rand = torch.rand(n, Sc + 1, device=device)
            nc = noise(Sc)
            u = torch.empty(n, Sc + 1, device=device).uniform_(to=(1 / (Sc + 1) - _EPS32))
            return t_rand, u, nc, noise(Sc)
        t_rand = torch.rand(n, Sc, device=device) if self.perturb else None
        nc = noise(Sc)
        u = torch.rand(n, Nf).to(device) if self.perturb else None
        return t_rand, u, nc, noise(Sc + Nf)

    def _draws(self, n, device):
        """mip mode jitters samples even at inference (``randomized`` is never overridden,
        renderer.py:211-221).  With rng='torch' the two draws are made with the same calls, shapes
        and order as the reference (render_utils.py:444, 483-485), so a seeded run sees the same
        numbers; rng='philox' generates them inside the kernels instead."""
        if self.embed_type != "mip":
            return None, None, L.NMR_RNG_NONE, 0
        if self.rng == "philox":
            seed = int(torch.randint(0, 2**62, (1,)).item())
            return None, None, L.NMR_RNG_PHILOX, seed
        if self.rng == "none":
            return None, None, L.NMR_RNG_NONE, 0
        S1 = self._nets()[2] + 1
        t_rand = torch.rand(n, S1, device=device)
        u_jit = torch.empty(n, S1, device=device).uniform_(to=(1 / S1 - _EPS32))
        return t_rand, u_jit, L.NMR_RNG_NONE, 0

    # ------------------------------------------------------------------------------------
    def forward_nerf(self, model, pts, viewdirs=None, ret_pfeat=False, app_emb=None, validation=False):
        """Reference renderer.py:119-180: encode + evaluate one net on explicit sample points."""
        if self.embed_type == "mip":
            mean, var = pts
            n, m, _ = mean.shape
            x = self.xyz_encoder(mean.reshape(-1, 3), y=var.reshape(-1, 3))[0]
        else:
            n, m, _ = pts.shape
            x = self.xyz_encoder(pts.reshape(-1, 3))
        cols = [x]
        if self.use_viewdirs:
            cols.append(self.dirs_encoder(viewdirs).repeat_interleave(m, dim=0))
        if app_emb is not None:
            cols.append(app_emb.reshape(n, -1).repeat_interleave(m, dim=0))
        out = model.forward(torch.cat(cols, -1), ret_pfeat=ret_pfeat, val=validation)
        if ret_pfeat:
            raw, feats = out
            return raw.reshape(n, m, -1), feats.reshape(n, m, -1), None
        return out.reshape(n, m, -1)

    def render_rays(self, rays, ray_id=None, validation=False, _intermediates=False, _unnorm=None, _out_ptrs=None):
        """Hierarchical coarse->fine render of a ray batch (reference renderer.py:182-295).  ``_out_ptrs``: raw device
        addresses (e.g. peer memory of the matcher's rank, dist.PeerBuffers) that replace the freshly allocated tensors of
        the named outputs; those keys are then absent from the returned dict."""
        ray_mask = None
        if self.pfeat_mask is not None and self.ret_pfeat:
            # nerf_trainer.py:45-47, renderer.py:141-146, 250-255: features are returned for the masked rays only.  The reference
            # applies the mask inside NeRF.forward only when the tap is the last trunk layer (models/nerf.py:137-141); with
            # stop_layer >= 0 its own shapes no longer match, so that combination is refused here as well.
            if self._nets()[1].stop_layer >= 0:
                raise NotImplementedError("pfeat_mask with stop_layer >= 0 fails in the reference itself (models/nerf.py:137-141)")
            ray_mask = self.pfeat_mask.flatten().to(rays.device)
            if ray_mask.numel() != rays.shape[0]:
                raise ValueError(f"pfeat_mask has {ray_mask.numel()} entries for {rays.shape[0]} rays")
        if rays.device.type != "cuda":
            raise RuntimeError("NerfRenderer (nerfmatch_b200) renders CUDA ray tensors only: there is no CPU fallback")
        rays = rays.detach().float().contiguous()
        n, ncols = rays.shape
        eng = self.engine(rays.device)
        if self.appearance_embedding and ray_id is not None:
            ray_id = ray_id.detach().to(device=rays.device, dtype=torch.int64).contiguous()
        else:
            ray_id = None if not self.appearance_embedding else torch.ones(n, dtype=torch.int64, device=rays.device)
        noise_c = noise_f = None
        if self.training:                   # perturb / noise_std act in training mode only (renderer.py:217, 242)
            t_rand, u_jit, noise_c, noise_f = self._train_draws(n, rays.device)
            rng_mode, seed = L.NMR_RNG_NONE, 0
        else:
            t_rand, u_jit, rng_mode, seed = self._draws(n, rays.device)
        outs = eng.alloc_outputs(n, self.ret_pfeat, _unnorm is not None, intermediates=_intermediates, training=not validation)
        if _out_ptrs:
            outs.update(_out_ptrs)
        args, keep = eng.make_args(n, ncols, rays=rays, ray_id=ray_id, t_rand=t_rand, u_jitter=u_jit, flags=self._flags(),
                                   rng_mode=rng_mode, seed=seed, unnorm=_unnorm, outs=outs, sigma_noise_coarse=noise_c, sigma_noise_fine=noise_f)
        eng.render(args)
        outs = {k: v for k, v in outs.items() if not isinstance(v, int)}
        preds = {}

        def put(k):                          # (outputs redirected to raw addresses are not returned)
            if k in outs:
                preds[k] = outs[k][ray_mask] if (ray_mask is not None and k.startswith("feat_")) else outs[k]

        for key in ("coarse", "fine"):      # key order of the reference dict
            if self.ret_pfeat:
                put(f"feat_{key}")
            if validation:                  # pts_* only under validation (renderer.py:276-281) ...
                put(f"pts_{key}")
            elif key == "fine":             # ... s_fine / weights_fine only without it (:283-286)
                put("s_fine")
                put("weights_fine")
            put(f"rgb_{key}")
            put(f"depth_{key}")
        for k, v in outs.items():
            if k not in preds and (_intermediates or k == "pt3d"):
                preds[k] = v
        return preds

    def forward(self, rays, step=0, ray_id=None, validation=False):
        return self.render_rays(rays, ray_id, validation=validation)

    def predict(self, rays, w, h, out_raw=False, ray_id=None):
        self.set_training_mode(False)
        preds = self.forward(rays, validation=True, ray_id=ray_id)
        if out_raw:
            return preds
        for k in ("rgb_coarse", "depth_coarse", "rgb_fine", "depth_fine"):
            v = preds[k]
            if h * w == v.shape[0]:
                preds[k] = v.reshape(h, w, v.shape[1] if v.dim() > 1 else 1)
        return preds

    def render_novel_view(self, img_hw, K, c2w, unnorm_scene, device, downsample=8):
        """Reference renderer.py:315-333, with the rays generated on the device instead of on the
        host: only the pose crosses the bus."""
        self.ret_pfeat = True
        self.set_training_mode(False)
        H, W = img_hw
        if isinstance(unnorm_scene, np.ndarray):
            unnorm_scene = torch.from_numpy(unnorm_scene)
        unnorm_scene = unnorm_scene.detach().float().cpu()
        c2w_n = unnorm_scene.inverse() @ torch.as_tensor(c2w).detach().float().cpu()
        device = torch.device(device)
        eng = self.engine(device)
        h, w = len(range(downsample // 2, H, downsample)), len(range(downsample // 2, W, downsample))
        n = h * w
        t_rand, u_jit, rng_mode, seed = self._draws(n, device)
        ray_id = torch.ones(n, dtype=torch.int64, device=device) if self.appearance_embedding else None
        outs = eng.alloc_outputs(n, True, True)
        args, keep = eng.make_args(n, 12 if self.embed_type == "mip" else 11, ray_id=ray_id, t_rand=t_rand, u_jitter=u_jit,
                                   flags=self._flags(), rng_mode=rng_mode, seed=seed, unnorm=unnorm_scene, outs=outs)
        eng.render_view(torch.as_tensor(K).detach().float().cpu(), c2w_n, H, W, downsample, args)
        im = outs["rgb_fine"]
        if (W // downsample) * (H // downsample) == n:
            im = im.reshape(H // downsample, W // downsample, -1)
        return dict(im_pred=im, pt3d=outs["pt3d"], pt_feat=outs["feat_fine"])
