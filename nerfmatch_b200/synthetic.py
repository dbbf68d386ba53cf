"""Deterministic, torch-free synthetic weights and poses for benchmarks, smoke runs and parity fixtures.

There are no checkpoints offline, so every parity test, golden fixture and
benchmark uses random-init weights.  To keep fixtures small the weights are NOT stored:
they are regenerated from a seed with a counter-based hash (splitmix64), which is stable
across numpy / torch versions and machines.  Distribution follows ``nn.Linear``'s default
init, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias (reference model:
nerfmatch/nerf/models/nerf.py:44-64); key names are the reference ``state_dict`` keys
("Checkpoint / resume").
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, stream: int, count: int) -> np.ndarray:
    """count floats in [0,1) (24-bit mantissa) from (seed, stream, index)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed * 0x1000003 + stream], np.uint64))[0]
        idx = np.arange(count, dtype=np.uint64) + base
        bits = _splitmix64(idx)
    return ((bits >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def make_state_dict(
    seed: int,
    embed_type: str = "mip",
    xyz_num_freqs: int = 15,
    dirs_num_freqs: int = 4,
    hid: int = 256,
    layer_num: int = 8,
    skips=(4,),
    app_dim: int = 0,
    num_frames: int = 0,
    sigma_gain: float = 1.0,
    sigma_bias: float = 0.0,
) -> Dict[str, np.ndarray]:
    """Reference-keyed state_dict (numpy fp32).  ``sigma_gain`` / ``sigma_bias`` (scalar or
    (coarse, fine) pair) rescale ``alpha_linear``: raw default-init densities are ~-0.03 +- 0.01,
    i.e. ReLU kills every weight, so the fixtures calibrate them to a smooth or a peaky field
    (the latter exercises CDF inversion away from the near-uniform case)."""
    xyz_dim = 6 * xyz_num_freqs + (0 if embed_type == "mip" else 3)
    dirs_dim = 6 * dirs_num_freqs + 3
    sd: Dict[str, np.ndarray] = {}
    stream = [0]

    def lin(name: str, fan_out: int, fan_in: int):
        bound = 1.0 / np.sqrt(fan_in)
        stream[0] += 1
        w = uniform01(seed, stream[0], fan_out * fan_in).reshape(fan_out, fan_in)
        stream[0] += 1
        b = uniform01(seed, stream[0], fan_out)
        sd[name + ".weight"] = ((w * 2 - 1) * bound).astype(np.float32)
        sd[name + ".bias"] = ((b * 2 - 1) * bound).astype(np.float32)

    gains = sigma_gain if isinstance(sigma_gain, (tuple, list)) else (sigma_gain, sigma_gain)
    biases = sigma_bias if isinstance(sigma_bias, (tuple, list)) else (sigma_bias, sigma_bias)
    for ni, net in enumerate(("nerf_coarse", "nerf_fine")):
        for i in range(layer_num):
            if i == 0:
                fan_in = xyz_dim
            elif (i - 1) in skips:
                fan_in = hid + xyz_dim
            else:
                fan_in = hid
            lin(f"{net}.pts_linears.{i}", hid, fan_in)
        lin(f"{net}.views_linears.0", hid // 2, dirs_dim + hid + app_dim)
        lin(f"{net}.feature_linear", hid, hid)
        lin(f"{net}.alpha_linear", 1, hid)
        lin(f"{net}.rgb_linear", 3, hid // 2)
        sd[f"{net}.alpha_linear.weight"] = (sd[f"{net}.alpha_linear.weight"] * np.float32(gains[ni])).astype(np.float32)
        sd[f"{net}.alpha_linear.bias"] = (
            sd[f"{net}.alpha_linear.bias"] * np.float32(gains[ni]) + np.float32(biases[ni])
        ).astype(np.float32)
    if app_dim > 0:
        stream[0] += 1
        e = uniform01(seed, stream[0], num_frames * app_dim).reshape(num_frames, app_dim)
        sd["embedding_a.weight"] = ((e * 2 - 1) * 1.5).astype(np.float32)
    if embed_type == "mip":
        sd["xyz_encoder.scales"] = np.array([2**i for i in range(xyz_num_freqs)], np.int64)
        sd["dirs_encoder.scales"] = np.array([2**i for i in range(dirs_num_freqs)], np.int64)
    return sd


def make_pose(seed: int = 0):
    """A normalised camera inside the unit sphere: small seeded rotation, t=(0.1,0.05,-0.2)
    (the survey's probe pose)."""
    r = uniform01(seed, 991, 3).astype(np.float64) * 0.3 - 0.15
    cx, cy, cz = np.cos(r)
    sx, sy, sz = np.sin(r)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    c2w = np.eye(4)
    c2w[:3, :3] = Rz @ Ry @ Rx
    c2w[:3, 3] = [0.1, 0.05, -0.2]
    return c2w.astype(np.float32)


K_7SCENES = np.array([[525.0, 0, 320.0], [0, 525.0, 240.0], [0, 0, 1.0]], np.float32)
K_CAMBRIDGE = np.array([[740.0, 0, 427.0], [0, 740.0, 240.0], [0, 0, 1.0]], np.float32)


def make_unnorm_scene(scale: float = 3.0, t=(0.4, -0.2, 1.1)) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] *= np.float32(scale)
    m[:3, 3] = np.asarray(t, np.float32)
    return m
