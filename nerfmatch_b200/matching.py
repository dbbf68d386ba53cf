"""NeRFMatch-Mini coarse matching on the device: the consumer of the rendered 3D features.

Mirrors ``NeRFMatcherCoarse.coarse_matching`` (nerfmatch/nerfmatch_coarse_trainer.py:223-234) followed by
``extract_mutual_matches`` (nerfmatch/modules/extract_matches.py:13-40, inference branch) for batch size 1; the image
features come from the caller's backbone, the point features straight from ``render_novel_view``'s ``pt_feat``."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .engine import Engine, EngineConfig, _ptr

_engines = {}


def _engine(device) -> Engine:
    key = str(torch.device(device))
    if key not in _engines:
        _engines[key] = Engine(EngineConfig(), device)
    return _engines[key]


def coarse_match(im_feat: torch.Tensor, pt_feat: torch.Tensor, temperature: float = 10.0, temp_type: str = "mul",
                 im_mask: Optional[torch.Tensor] = None, pt_mask: Optional[torch.Tensor] = None, mutual: bool = False,
                 match_thres: float = 0.0, return_conf: bool = True):
    """im_feat [1,m,d] or [m,d], pt_feat [1,n,d] or [n,d] (CUDA fp32) -> dict(conf_matrix [1,m,n], match_ids=(b,i,j), mconf, pred_num)."""
    if im_feat.device.type != "cuda":
        raise RuntimeError("coarse_match runs on CUDA tensors only (no CPU fallback)")
    im = im_feat.detach().float().reshape(-1, im_feat.shape[-1]).contiguous()
    pt = pt_feat.detach().float().reshape(-1, pt_feat.shape[-1]).contiguous()
    m, d = im.shape
    n = pt.shape[0]
    dev = im.device
    eng = _engine(dev)
    conf = torch.empty(m, n, device=dev) if return_conf else None
    i_ids = torch.empty(m, dtype=torch.int64, device=dev)
    j_ids = torch.empty(m, dtype=torch.int64, device=dev)
    mconf = torch.empty(m, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    imk = None if im_mask is None else im_mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
    ptk = None if pt_mask is None else pt_mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
    with torch.cuda.device(dev):
        st = eng.lib.nmr_coarse_match(eng.h, _ptr(im), m, _ptr(pt), n, d, _ptr(imk), _ptr(ptk), float(temperature),
                                      1 if temp_type == "div" else 0, int(mutual), float(match_thres), _ptr(conf), _ptr(i_ids),
                                      _ptr(j_ids), _ptr(mconf), _ptr(count), eng.stream())
    eng.check(st, "nmr_coarse_match")
    k = int(count.item())
    out = dict(match_ids=(torch.zeros(k, dtype=torch.int64, device=dev), i_ids[:k], j_ids[:k]), mconf=mconf[:k], pred_num=k)
    if return_conf:
        out["conf_matrix"] = conf[None]
    return out
