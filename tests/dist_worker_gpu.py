"""Worker of tests/test_gpu_surfaces.py::test_render_views_sharded_two_gpus (one process per GPU, NCCL).
Rank 0 compares the gathered maps of `render_views_sharded` with the unsharded `render_novel_view` of every view."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

from nmr_helpers import build_renderer  # noqa: E402
from nerfmatch_b200.dist import render_views_sharded  # noqa: E402
from nerfmatch_b200.synthetic import K_7SCENES, make_pose, make_state_dict, make_unnorm_scene  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank()
    c = dict(embed="normal", nc=64, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False)
    sd = make_state_dict(91, embed_type="normal", sigma_gain=60.0, sigma_bias=(4.0, 4.0))
    ok = True
    for prec in ("fp32", "bf16"):
        ren = build_renderer(c, sd, precision=prec, device=dev)
        un = make_unnorm_scene()
        V = 5                                                      # 5 views x 1200 rays: shard boundaries fall inside views
        c2ws = torch.stack([torch.from_numpy(un @ make_pose(30 + v)) for v in range(V)])
        out = render_views_sharded(ren, (480, 640), torch.from_numpy(K_7SCENES), c2ws, un, downsample=16, dst=0)
        if rank == 0:
            for v in range(V):
                ref = ren.render_novel_view((480, 640), torch.from_numpy(K_7SCENES), c2ws[v], un, dev, downsample=16)
                for a, b in (("pt_feat", "pt_feat"), ("pt3d", "pt3d"), ("im_pred", "im_pred")):
                    x, y = out[a][v], ref[b].reshape(out[a][v].shape)
                    if prec == "fp32" or a != "pt_feat":
                        same = torch.equal(x, y)
                    else:      # bf16: features are reduced per 128-point tile, whose boundaries depend on the shard's first ray
                        same = bool((x - y).abs().max() <= 2e-5 * y.abs().max())
                    if not same:
                        ok = False
                        print(f"MISMATCH {prec} view {v} {a}: max |d| = {(x - y).abs().max().item():.3e}")
        else:
            assert out is None
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and ok:
        print("SHARDED_OK")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
