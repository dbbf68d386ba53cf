"""Multi-GPU plumbing for the render path: one process per GPU (torch.distributed, NCCL over
NVLink / NVSwitch on the box, gloo in the CPU tests).

Rays are independent, so the path shards with NO data-path collective: the global ray
index space - (view, pixel) flattened - is cut into contiguous, tile-aligned blocks, one per rank; every
rank holds a replica of both nets' packed weights.  The only exchange is at the end: the rendered
feature map (``feat_fine`` [n_r,256]) and 3D points ([n_r,3]) are gathered to the matcher's rank
(``dist.gather``: NCCL send/recv pairs; sub-millisecond for matcher-resolution views).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 128  # MLP M-tile: shard boundaries are aligned to it so no rank renders a ragged tile but the last


def shard_range(n: int, world: int, rank: int, align: int = TILE) -> Tuple[int, int]:
    """Contiguous block [start, end) of ``n`` rays owned by ``rank``: ceil(n/world) rounded up to ``align``."""
    if n <= 0:
        return 0, 0
    per = -(-n // world)
    per = -(-per // align) * align
    start = min(n, rank * per)
    return start, min(n, start + per)


def shard_sizes(n: int, world: int, align: int = TILE) -> List[int]:
    return [e - s for s, e in (shard_range(n, world, r, align) for r in range(world))]


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0, group=None, align: int = TILE) -> Optional[torch.Tensor]:
    """Gather per-rank row blocks (sharded with ``shard_range``) to ``dst``; returns [n_total, ...] there, None elsewhere.
    Blocks are padded to the common shard size so the exchange is one fixed-size gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world, align)
    per = max(sizes) if sizes else 0
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    if per == 0:   # nothing to exchange: every shard is empty
        return local.new_empty((0,) + tuple(local.shape[1:])) if rank == dst else None
    if local.dim() == 1:
        local = local.reshape(-1, 1)
    send = local
    if local.shape[0] != per:
        send = local.new_zeros((per,) + tuple(local.shape[1:]))
        send[: local.shape[0]] = local
    send = send.contiguous()
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def render_rays_sharded(render_fn: Callable[[torch.Tensor], Dict[str, torch.Tensor]], rays: torch.Tensor,
                        keys: Sequence[str] = ("feat_fine", "pts_fine"), dst: int = 0, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """Every rank holds the same ``rays`` [n, c] (or at least its own block); each renders its block with
    ``render_fn`` (e.g. ``lambda r: renderer.predict(r, 1, 1)``) and the requested outputs are gathered to ``dst``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = rays.shape[0]
    s, e = shard_range(n, world, rank)
    preds = render_fn(rays[s:e])
    out = {}
    for k in keys:
        t = preds[k]
        g = gather_rows(t.reshape(e - s, t.numel() // max(e - s, 1)) if e > s else t.reshape(0, *t.shape[1:]), n, dst=dst, group=group)
        if rank == dst:
            out[k] = g
    return out if rank == dst else None


def render_views_sharded(renderer, img_hw, K, c2ws: torch.Tensor, unnorm_scene, downsample: int = 8, dst: int = 0, group=None):
    """BASELINE config 4: a batch of V query poses, rays of all views flattened and sharded across the ranks,
    ``pt_feat`` / ``pt3d`` / ``im_pred`` gathered to the matcher rank as [V, h*w, .] tensors.
    Poses are world poses (as passed to ``render_novel_view``); ray generation happens on each rank's device."""
    import numpy as np

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    device = torch.device("cuda", torch.cuda.current_device())
    H, W = img_hw
    h, w = len(range(downsample // 2, H, downsample)), len(range(downsample // 2, W, downsample))
    per_view = h * w
    V = c2ws.shape[0]
    n = V * per_view
    s, e = shard_range(n, world, rank)
    un = torch.as_tensor(np.asarray(unnorm_scene) if not isinstance(unnorm_scene, torch.Tensor) else unnorm_scene).float().cpu()
    eng = renderer.engine(device)
    blocks = []
    for v in range(s // per_view, (e + per_view - 1) // per_view if e > s else 0):
        c2w_n = un.inverse() @ c2ws[v].float().cpu()
        rays_v, _, _ = eng.raygen(torch.as_tensor(K).float().cpu(), c2w_n, H, W, downsample)
        lo, hi = max(s, v * per_view) - v * per_view, min(e, (v + 1) * per_view) - v * per_view
        blocks.append(rays_v[lo:hi])
    ncols = 12 if renderer.embed_type == "mip" else 11
    rays = torch.cat(blocks, 0) if blocks else torch.empty(0, ncols, device=device)
    renderer.ret_pfeat = True
    preds = renderer.render_rays(rays, validation=True, _unnorm=un)
    out = {}
    for name, key in (("pt_feat", "feat_fine"), ("pt3d", "pt3d"), ("im_pred", "rgb_fine")):
        t = preds[key]
        g = gather_rows(t.reshape(e - s, t.numel() // max(e - s, 1)) if e > s else t.reshape(0, *t.shape[1:]), n, dst=dst, group=group)
        if rank == dst:
            out[name] = g.reshape(V, per_view, g.shape[-1])
    return out if rank == dst else None
