"""Batched scene-feature cache writer.

The reference renders one frame per iteration of a host loop and ``np.save``s a pickled dict per frame
(``NerfEvaluator.cache_scene_pts``, nerfmatch/nerf_evaluator.py:308-372); the NeRFMatch datasets read it back with
``load_frame_3d`` (nerfmatch/datasets/data_loading.py:36-80).  Here several views are rendered per launch (rays of
all views concatenated, generated on the device from the poses), results come back through pinned buffers on a
side stream, and a small thread pool writes the files - in exactly the reference's format:

    {img_idx}.npy  ->  dict(pt3d [n,3] f32 un-normalised, unnorm_scene [4,4] f32, pt_feat [n,256] f32,
                            pt_color [n,3] f32 clamped to [0,1])
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch


def _save(path: str, item: Dict[str, np.ndarray]):
    np.save(path, item)   # a dict -> pickled 0-d object array, like the reference's np.save(scene_dir / f"{img_idx}.npy", scene_pts)


def write_frame(scene_dir: str, img_idx: str, pt3d, unnorm_scene, pt_feat, pt_color, **extra) -> str:
    """One frame of the scene cache in the reference's format (nerf_evaluator.py:358-372): ``{img_idx}.npy`` holding the
    pickled dict(pt3d [n,3] f32, unnorm_scene [4,4] f32, pt_feat [n,C] f32, pt_color [n,3] f32 in [0,1]); ``extra`` carries
    the optional keys the reference adds for some datasets (sky_mask, mask).  Pure numpy: no device involved."""
    item = dict(pt3d=np.ascontiguousarray(pt3d, np.float32), unnorm_scene=np.ascontiguousarray(unnorm_scene, np.float32),
                pt_feat=np.ascontiguousarray(pt_feat, np.float32), pt_color=np.clip(np.ascontiguousarray(pt_color, np.float32), 0.0, 1.0))
    item.update(extra)
    path = os.path.join(scene_dir, f"{img_idx}.npy")
    _save(path, item)
    return path


def cache_scene_pts(renderer, frames: Sequence[Dict], scene_dir: str, img_hw, unnorm_scene, downsample: int = 8,
                    feat_comb: str = "lin", views_per_batch: int = 8, device=None, writer_threads: int = 2,
                    rank: int = 0, world_size: int = 1) -> List[str]:
    """Render and cache ``frames`` (each: dict(img_idx=str, K=[3,3], c2w=[4,4] world pose[, ts=int])).  ``ts`` is the frame's
    appearance-embedding id (the reference passes ``batch['ts']`` as ``ray_id``, nerf_evaluator.py:343-349); without it the
    renderer's default id 1 is used, as ``render_novel_view`` does.

    With ``world_size`` > 1 the frames are dealt round-robin to the ranks (no collective: every rank writes its own
    files).  Returns the paths written by this rank."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    os.makedirs(scene_dir, exist_ok=True)
    renderer.ret_pfeat = True
    renderer.feat_comb = feat_comb
    renderer.set_training_mode(False)
    un = torch.as_tensor(np.asarray(unnorm_scene) if not isinstance(unnorm_scene, torch.Tensor) else unnorm_scene).float().cpu()
    un_inv = un.inverse()
    eng = renderer.engine(device)
    H, W = img_hw
    mine = [f for i, f in enumerate(frames) if i % world_size == rank]
    copy_stream = torch.cuda.Stream(device)
    pool = ThreadPoolExecutor(max(1, writer_threads))
    futures, paths = [], []
    pending = None   # (event, host tensors, frames of the batch, rays per view)

    def flush(p):
        ev, host, batch, per_view = p
        ev.synchronize()
        for v, f in enumerate(batch):
            sl = slice(v * per_view, (v + 1) * per_view)
            paths.append(os.path.join(scene_dir, f"{f['img_idx']}.npy"))
            futures.append(pool.submit(write_frame, scene_dir, f["img_idx"], host["pt3d"][sl].numpy().copy(), un.numpy().copy(),
                                       host["feat_fine"][sl].numpy().copy(), host["rgb_fine"][sl].numpy().copy()))

    for b0 in range(0, len(mine), views_per_batch):
        batch = mine[b0:b0 + views_per_batch]
        rays, ids = [], []
        for f in batch:
            c2w_n = un_inv @ torch.as_tensor(np.asarray(f["c2w"])).float().cpu()
            r, (h, w), _ = eng.raygen(torch.as_tensor(np.asarray(f["K"])).float().cpu(), c2w_n, H, W, downsample)
            rays.append(r)
            ids.append(torch.full((r.shape[0],), int(f.get("ts", 1)), dtype=torch.int64, device=device))
        per_view = rays[0].shape[0]
        ray_id = torch.cat(ids, 0) if getattr(renderer, "appearance_embedding", False) else None
        preds = renderer.render_rays(torch.cat(rays, 0), ray_id=ray_id, validation=True, _unnorm=un)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(device))
        host = {}
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            for k in ("pt3d", "feat_fine", "rgb_fine"):
                host[k] = torch.empty(preds[k].shape, dtype=preds[k].dtype, pin_memory=True)
                host[k].copy_(preds[k], non_blocking=True)
                preds[k].record_stream(copy_stream)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        if pending is not None:
            flush(pending)          # write batch i while batch i+1 renders
        pending = (ev, host, batch, per_view)
    if pending is not None:
        flush(pending)
    for fu in futures:
        fu.result()
    pool.shutdown()
    return paths


def frame_cache_name(file_path: str) -> str:
    """File stem the reference derives from a frame's ``file_path`` (datasets/data_loading.py:39-40)."""
    return file_path.replace("/", "_").replace(".color", "").replace(".png", "")


def load_frame_3d(frame, scene_dir, use_msk=None, return_pose=False):
    """Reader with the reference's signature and semantics (datasets/data_loading.py:36-80):
    -> pt3d, pt_feat, mask, unnorm_scene[, c2w].  ``frame`` is the dataset's frame dict (``file_path`` names the cache file)."""
    scene_pts = np.load(os.path.join(scene_dir, f"{frame_cache_name(frame['file_path'])}.npy"), allow_pickle=True).item()
    pt3d = scene_pts["pt3d"]
    unnorm_scene = scene_pts["unnorm_scene"]
    c2w = torch.from_numpy(unnorm_scene @ scene_pts["cam2scene"]) if "cam2scene" in scene_pts else None
    mask = np.ones(len(pt3d)).astype(np.bool_)
    if "pt_mask" in scene_pts and use_msk:
        src = {"sky": lambda: scene_pts["sky_mask"][0], "corr": lambda: scene_pts["corr_mask"]}.get(use_msk, lambda: scene_pts["pt_mask"][0])()
        mask = (1 - src.reshape(-1)).astype(np.bool_)
    if return_pose:
        return pt3d, scene_pts["pt_feat"], mask, unnorm_scene, c2w
    return pt3d, scene_pts["pt_feat"], mask, unnorm_scene
