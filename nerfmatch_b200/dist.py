"""Multi-GPU plumbing for the render path: one process per GPU (torch.distributed, NCCL over
NVLink / NVSwitch on the box, gloo in the CPU tests).

Rays are independent, so the path shards with NO data-path collective: the global ray
index space - (view, pixel) flattened - is cut into contiguous, tile-aligned blocks, one per rank; every
rank holds a replica of both nets' packed weights.  The only exchange is at the end: the rendered
feature map (``feat_fine`` [n_r,256]) and 3D points ([n_r,3]) go to the matcher's rank.  Two forms:

* ``PeerBuffers`` (the product path on a box): the matcher rank allocates the gathered maps once and exports
  them over CUDA IPC; every other rank maps them (NVLink / NVSwitch peer memory) and passes the mapped addresses as
  the OUTPUT pointers of its render - the compositing kernel's own stores deliver the rows, chunk by chunk, while later
  chunks are still in the MLP.  No collective follows the render; a release / acquire flag per producer orders the
  matcher rank's stream behind the stores.
* ``gather_rows`` (``dist.gather``: NCCL send / recv after the render) - the portable form, also what the gloo tests
  exercise on CPU.
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


class _CudaView:
    """__cuda_array_interface__ holder: lets torch alias library-owned device memory (no copy)."""

    def __init__(self, ptr: int, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
        self._owner = owner


class PeerBuffers:
    """Gathered output maps that live on rank ``dst`` and are written by every rank's render kernels directly.

    ``spec`` maps an output key of nmr_render_args (``feat_fine``, ``pt3d``, ``rgb_fine`` ...) to (total rows, columns);
    rank r writes the rows of its shard (``shard_range``: the row offset IS the global ray index).  Protocol per step:
    ``begin()`` (a stream-ordered barrier: the matcher has consumed the previous step) -> render with ``out_ptrs(row0)`` as
    outputs -> ``publish()`` -> on ``dst``: ``wait()`` then read ``tensor(key)``."""

    def __init__(self, eng, spec: Dict[str, Tuple[int, int]], dst: int = 0, group=None):
        self.eng, self.group, self.dst = eng, group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.spec = dict(spec)
        self.off, off = {}, 256                                   # [0, 256): one uint32 step flag per producer
        for k, (rows, cols) in self.spec.items():
            self.off[k] = off
            off += -(-rows * cols * 4 // 256) * 256
        self.nbytes = off
        handle = None
        self.base = 0
        if self.rank == dst:
            self.base, handle = eng.peer_alloc(self.nbytes)
        box = [handle]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        if self.rank != dst:
            self.base = eng.peer_open(box[0])
        self.step = 0
        self._sync = torch.zeros(1, device=eng.device)

    def out_ptrs(self, row0: int) -> Dict[str, int]:
        return {k: self.base + self.off[k] + row0 * self.spec[k][1] * 4 for k in self.spec}

    def begin(self):
        """All ranks: the matcher's stream has passed every read of the previous step (stream-ordered tiny all-reduce)."""
        if self.step > 0:
            dist.all_reduce(self._sync, group=self.group)

    def publish(self):
        self.step += 1
        self.eng.peer_signal(self.base + 4 * self.rank, self.step)

    def wait(self):
        assert self.rank == self.dst
        self.eng.peer_wait(self.base, self.world, self.step)

    def tensor(self, key: str) -> torch.Tensor:
        """dst only: [rows, cols] view of the gathered map (valid until the next step's ``begin``)."""
        assert self.rank == self.dst
        return torch.as_tensor(_CudaView(self.base + self.off[key], self.spec[key], self), device=self.eng.device)

    def check(self):
        """dst, after synchronising: raises if a producer's flag did not arrive within the device-side timeout."""
        late = self.eng.peer_wait_status()
        if late:
            raise RuntimeError(f"peer exchange: rank {late - 1} did not deliver step {self.step} in time")

    def close(self):
        if self.base:
            torch.cuda.synchronize(self.eng.device)
            dist.barrier(group=self.group)
            (self.eng.peer_free if self.rank == self.dst else self.eng.peer_close)(self.base)
            self.base = 0


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


def render_views_sharded(renderer, img_hw, K, c2ws: torch.Tensor, unnorm_scene, downsample: int = 8, dst: int = 0, group=None,
                         exchange: str = "auto"):
    """BASELINE config 4: a batch of V query poses, rays of all views flattened and sharded across the ranks,
    ``pt_feat`` / ``pt3d`` / ``im_pred`` delivered to the matcher rank as [V, h*w, .] tensors.
    Poses are world poses (as passed to ``render_novel_view``); ray generation happens on each rank's device.
    ``exchange``: "peer" - the render kernels store into the matcher rank's memory (PeerBuffers; the returned tensors alias
    that buffer and stay valid until the next call), "nccl" - dist.gather after the render, "auto" - peer on NCCL groups."""
    import numpy as np

    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_initialized() else (1, 0)
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
    if exchange == "auto":
        exchange = "peer" if (world > 1 and dist.get_backend(group) == "nccl") else "nccl"
    if exchange == "peer" and world > 1:
        cache = renderer.__dict__.setdefault("_peer_buffers", {})
        key = (n, world, dst, str(device))
        if key not in cache:
            cache[key] = PeerBuffers(eng, {"feat_fine": (n, 256), "pt3d": (n, 3), "rgb_fine": (n, 3)}, dst=dst, group=group)
        pb = cache[key]
        pb.begin()
        renderer.render_rays(rays, validation=True, _unnorm=un, _out_ptrs=pb.out_ptrs(s))
        pb.publish()
        if rank != dst:
            return None
        pb.wait()
        return {name: pb.tensor(k).reshape(V, per_view, -1) for name, k in (("pt_feat", "feat_fine"), ("pt3d", "pt3d"), ("im_pred", "rgb_fine"))}
    preds = renderer.render_rays(rays, validation=True, _unnorm=un)
    if world == 1:
        return {name: preds[key].reshape(V, per_view, -1) for name, key in (("pt_feat", "feat_fine"), ("pt3d", "pt3d"), ("im_pred", "rgb_fine"))}
    out = {}
    for name, key in (("pt_feat", "feat_fine"), ("pt3d", "pt3d"), ("im_pred", "rgb_fine")):
        t = preds[key]
        g = gather_rows(t.reshape(e - s, t.numel() // max(e - s, 1)) if e > s else t.reshape(0, *t.shape[1:]), n, dst=dst, group=group)
        if rank == dst:
            out[name] = g.reshape(V, per_view, g.shape[-1])
    return out if rank == dst else None
