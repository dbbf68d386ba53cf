"""``NeRF`` with the reference's parameter layout (nerfmatch/nerf/models/nerf.py:14-64): the
nn.Linear parameters are the source of truth for checkpoints; ``forward`` evaluates them with the
CUDA MLP through the C ABI (nmr_stage_mlp)."""
from __future__ import annotations

from argparse import Namespace

import torch
import torch.nn as nn

from .autograd import MlpFn, wants_grad
from .engine import Engine, EngineConfig

_DEFAULTS = dict(layer_num=8, hid_dim=256, xyz_dim=3, dirs_dim=3, app_dim=0, output_dim=4, skips=[4],
                 use_viewdirs=False, out_3d_pnt=False, out_add_ch=0, stop_layer=-1)


class NeRF(nn.Module):
    default_config = dict(_DEFAULTS)

    def __init__(self, config):
        super().__init__()
        conf = dict(_DEFAULTS)
        given = vars(config) if isinstance(config, Namespace) else dict(config)
        conf.update({k: v for k, v in given.items() if k in conf})
        self.layer_num, self.xyz_dim, self.dirs_dim = conf["layer_num"], conf["xyz_dim"], conf["dirs_dim"]
        self.app_dim, self.output_dim, self.skips = conf["app_dim"], conf["output_dim"], list(conf["skips"])
        self.use_viewdirs, self.out_3d_pnt, self.stop_layer = conf["use_viewdirs"], conf["out_3d_pnt"], conf["stop_layer"]
        hid = conf["hid_dim"]
        if not self.use_viewdirs or self.out_3d_pnt or hid != 256 or self.layer_num != 8 or self.skips != [4] or self.output_dim != 4:
            raise NotImplementedError(
                "nerfmatch_b200 builds the shipped NeRF only: 8x256 trunk, skips=[4], use_viewdirs=True, output_dim=4, no out_scr head"
            )
        self.pts_linears = nn.ModuleList(
            [nn.Linear(self.xyz_dim, hid)]
            + [nn.Linear(hid + (self.xyz_dim if i in self.skips else 0), hid) for i in range(self.layer_num - 1)]
        )
        self.views_linears = nn.ModuleList([nn.Linear(self.dirs_dim + hid + self.app_dim, hid // 2)])
        self.feature_linear = nn.Linear(hid, hid)
        self.alpha_linear = nn.Linear(hid, 1)
        self.rgb_linear = nn.Linear(hid // 2, self.output_dim - 1)
        self._engines = {}

    def net_state(self):
        return {k: v for k, v in self.state_dict().items()}

    def weights_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def forward(self, x, ret_pfeat=0, pfeat_mask=None, val=False):
        """x: [P, xyz+dirs+app] in the reference column order -> [P,4] (rgb, raw sigma) and, if
        ``ret_pfeat``, the tapped [P,256] activations.  Differentiable w.r.t. ``x`` (not the weights)."""
        if pfeat_mask is not None:
            raise NotImplementedError("pfeat_mask is a training-validation feature (out of scope)")
        if x.device.type != "cuda":
            raise RuntimeError("nerfmatch_b200.NeRF runs on CUDA tensors only (no CPU fallback)")
        key = str(x.device)
        sig = self.weights_signature()
        ent = self._engines.get(key)
        if ent is None or ent[1] != sig:
            F = (self.xyz_dim - (3 if self.xyz_dim % 6 else 0)) // 6
            cfg = EngineConfig(embed_type="normal" if self.xyz_dim % 6 else "mip", xyz_num_freqs=F,
                               dirs_num_freqs=(self.dirs_dim - 3) // 6, app_dim=self.app_dim,
                               precision=getattr(self, "precision", "fp32"))
            eng = ent[0] if ent is not None else Engine(cfg, x.device)
            eng.set_weights(0, self.net_state())
            eng.set_weights(1, self.net_state())
            ent = (eng, sig)
            self._engines[key] = ent
        if wants_grad(x):           # iNeRF refinement (nerfmatch_evaluator.py:395-404): d (raw, feat) / d x
            raw, feat = MlpFn.apply(x.reshape(-1, x.shape[-1]), ent[0], 0, self.stop_layer)
        else:
            xf = x.detach().reshape(-1, x.shape[-1]).float().contiguous()
            raw, feat = ent[0].stage_mlp(0, xf, self.stop_layer, want_feat=ret_pfeat > 0)
        raw = raw.reshape(*x.shape[:-1], 4)
        if ret_pfeat > 0:
            return raw, feat.reshape(*x.shape[:-1], 256)
        return raw
