"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard arithmetic and the gather protocol.
The render itself is replaced by a deterministic per-ray stub (the CUDA path cannot run here); what is
checked is that sharded + gathered == un-sharded for ragged sizes, empty shards and multi-column outputs."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from nerfmatch_b200.dist import gather_rows, render_rays_sharded, shard_range, shard_sizes  # noqa: E402


def test_shard_ranges_cover_and_align():
    for n in (0, 1, 127, 128, 129, 3600, 4800, 6420, 307200, 64 * 3600 + 5):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (s0, e0), (s1, e1) in zip(ranges, ranges[1:]):
                assert e0 == s1 and s0 <= e0
            for s, e in ranges[:-1]:
                assert e == s or (s % 128 == 0 and (e % 128 == 0 or e == n))
            assert sum(shard_sizes(n, world)) == n


def _stub_render(rays):
    # deterministic, per-ray independent stand-in for NerfRenderer.predict (CPU test only)
    f = torch.cat([rays.sum(-1, keepdim=True) * (i + 1) for i in range(8)], -1)
    return {"feat_fine": f, "pts_fine": rays[:, :3] * 2.0 + 1.0}


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rays = torch.arange(n * 11, dtype=torch.float32).reshape(n, 11) / 7.0
        out = render_rays_sharded(_stub_render, rays, dst=0)
        if rank == 0:
            ref = _stub_render(rays)
            assert torch.equal(out["feat_fine"], ref["feat_fine"]) and torch.equal(out["pts_fine"], ref["pts_fine"])
        else:
            assert out is None
        # direct gather of a ragged int tensor to a non-zero destination
        s, e = shard_range(n, world, rank)
        loc = torch.arange(s, e, dtype=torch.int64).reshape(-1, 1)
        g = gather_rows(loc, n, dst=world - 1)
        if rank == world - 1:
            assert torch.equal(g.flatten(), torch.arange(n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [0, 5, 128, 300, 3600])
def test_sharded_render_equals_unsharded_gloo(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
