"""Shared helpers for the parity tests."""
import ast
import os

import numpy as np

from oracle.nerf_oracle import RenderConfig
from oracle.weights import make_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


RENDER_CASES = [
    "mip64", "mip128_peaky", "normal_64_128", "normal_64_128_peaky",
    "mip64_app_max_white", "normal_app_stop5",
]


def case_setup(g):
    """golden render_* fixture -> (RenderConfig, state_dict, case dict)."""
    c = ast.literal_eval(str(g["case"]))
    cfg = RenderConfig(
        embed_type=c["embed"], num_pts_coarse=c["nc"], num_pts_fine=c["nf"],
        app_dim=16 if c["app"] else 0, stop_layer=c["stop_layer"], feat_comb=c["feat_comb"],
        white_bg=c["white_bg"],
    )
    sd = make_state_dict(
        c["seed"], embed_type=c["embed"], app_dim=cfg.app_dim, num_frames=7 if c["app"] else 0,
        sigma_gain=c["sigma_gain"], sigma_bias=c["sigma_bias"],
    )
    return cfg, sd, c


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def ray_rel_err(a, b):
    """per-ray relative L2 error of [n, C] (or [n]) arrays"""
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)


def cosine_rows(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    num = (a * b).sum(-1)
    den = np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1)
    return num / np.maximum(den, 1e-30)


from nerfmatch_b200.config import build_renderer as _build, make_config as ref_style_config  # noqa: E402,F401


def build_renderer(c, sd, precision="fp32", device="cuda:0", **kw):
    return _build(c, sd, precision=precision, device=device, **kw)
