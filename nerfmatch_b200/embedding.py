"""Encoder modules with the reference's names, constructor arguments and state_dict keys
(nerfmatch/nerf/embedding.py:13-84).  They hold no arithmetic of their own: ``forward`` runs the
CUDA encoding kernel through the C ABI (nmr_stage_encode)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .autograd import EncodeFn, wants_grad
from .engine import Engine, EngineConfig


class _EncoderBase(nn.Module):
    _embed_type = "normal"

    def __init__(self, num_freqs: int):
        super().__init__()
        self.num_freqs = num_freqs
        self._engines = {}

    def get_embedding_dim(self, in_dim: int) -> int:
        return 2 * in_dim * self.num_freqs + in_dim

    def _engine(self, device) -> Engine:
        key = str(device)
        if key not in self._engines:
            # a private, weight-less handle: both frequency counts set to ours, kind=1 path is used
            cfg = EngineConfig(embed_type=self._embed_type, xyz_num_freqs=min(self.num_freqs, 16),
                               dirs_num_freqs=min(self.num_freqs, 8))
            self._engines[key] = Engine(cfg, device)
        return self._engines[key]

    def _encode(self, x: torch.Tensor, var=None) -> torch.Tensor:
        if x.device.type != "cuda":
            raise RuntimeError("nerfmatch_b200 encoders run on CUDA tensors only (no CPU fallback)")
        shape = x.shape[:-1]
        vf = None if var is None else var.detach().reshape(-1, 3).float().contiguous()
        eng = self._engine(x.device)
        # the C ABI encodes with the handle's xyz slot (kind 0: up to 16 frequencies, IPE when a variance is given) or its
        # dirs slot (kind 1: up to 8); this private handle has both slots set to our frequency count
        kind = 0 if self.num_freqs > 8 or vf is not None else 1
        if wants_grad(x):           # iNeRF refinement: d enc / d x (the variances are constants there)
            out = EncodeFn.apply(x.reshape(-1, 3), vf, eng, kind)
        else:
            out = eng.stage_encode(kind, x.detach().reshape(-1, 3).float().contiguous(), vf)
        return out.reshape(*shape, out.shape[-1])


class FourierEmbedding(_EncoderBase):
    """(x, sin(2^k x), cos(2^k x), ...) - reference embedding.py:13-49 (logscale, scale=1)."""

    _embed_type = "normal"

    def __init__(self, num_freqs, logscale=True, scale=1.0):
        super().__init__(num_freqs)
        if not logscale or scale != 1.0:
            raise NotImplementedError("only logscale=True, scale=1.0 (the shipped configuration) is built")
        self.logscale, self.scale = logscale, scale
        self.freqs = 2 ** torch.linspace(0, num_freqs - 1, num_freqs)

    def forward(self, x, **kwargs):
        return self._encode(x)

    def __repr__(self):
        return f"FourierEmbedding(num_freqs={self.num_freqs}, logscale={self.logscale}, scale={self.scale})"


class PositionalEncodingMIP(_EncoderBase):
    """IPE for (mean, var) pairs, PE (input appended) otherwise - reference embedding.py:52-84.
    ``scales`` is kept as a frozen int64 Parameter because it is a checkpoint key."""

    _embed_type = "mip"

    def __init__(self, num_freqs, min_deg=0):
        super().__init__(num_freqs)
        if min_deg != 0:
            raise NotImplementedError("min_deg != 0 is not built")
        self.min_deg, self.max_deg = min_deg, num_freqs
        self.scales = nn.Parameter(torch.tensor([2**i for i in range(min_deg, num_freqs)]), requires_grad=False)

    def forward(self, x, y=None):
        if y is not None:
            # the reference also returns the encoded variance; no caller reads it (renderer.py:156)
            return self._encode(x, y), None
        return self._encode(x)
