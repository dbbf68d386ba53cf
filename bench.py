#!/usr/bin/env python
"""Benchmark of the scene-feature render hot path (BASELINE.json: rays/sec, feature render, 64+128 samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp16|fp32]

A "step" renders one full 480x640 view (307 200 rays, classic 64 coarse + 128 fine samples = 256 MLP
evaluations per ray, feature tap at layer 3, all reference outputs incl. feat_coarse) per GPU with
random-init weights (oracle/weights.py) and a synthetic pose.  Prints ONE JSON line (rank 0):

  value     rays/s of the whole job with the rays already resident in HBM (CUDA events, max over ranks)
  e2e       rays/s through the C-ABI host entry (nmr_render_host): pinned host rays in, host outputs back
  roofline  the MLP kernels' achieved TFLOP/s (algorithmic FLOPs / CUDA-event time) vs the measured bf16 peak
  cpu_baseline  the oracle port on the host cores, on a bounded sample of the same workload (rank 0, N=1)

`--impl reference` times the CPU implementation only (the oracle port: the reference is Python and cannot
travel to the GPU box) on the same config / metric / unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 480, 640
METRIC = "rays/sec (feature render, 64+128 samples)"
FLOP_PER_EVAL = {"normal": 1217536, "mip": 1214464}   # 2 x weights-only MACs, un-padded


def workload(embed="normal"):
    return dict(embed=embed, nc=64, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=False)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores (bounded sample)
# --------------------------------------------------------------------------------------------
def cpu_render_rate(embed, n_sample, seed=0):
    from concurrent.futures import ThreadPoolExecutor

    from threadpoolctl import threadpool_limits

    from oracle import nerf_oracle as O
    from oracle.weights import K_7SCENES, make_pose, make_state_dict, uniform01

    cores = os.cpu_count() or 1
    cfg = O.RenderConfig(embed_type=embed, num_pts_coarse=64, num_pts_fine=128, stop_layer=3)
    sd = make_state_dict(seed, embed_type=embed, sigma_gain=120.0, sigma_bias=(8.0, 8.0))
    ds = max(1, int(np.sqrt(H * W / n_sample)))
    rays = O.sample_nerf_rays(H, W, K_7SCENES, make_pose(seed), ds=ds, embed_type=embed)[:n_sample]
    n = rays.shape[0]
    tr = uj = None
    if embed == "mip":
        tr = uniform01(seed, 5, n * 65).reshape(n, 65)
        uj = (uniform01(seed, 6, n * 65).reshape(n, 65) * (1 / 65 - 1.2e-7)).astype(np.float32)
    ch = 64

    def work(i):
        return O.render_rays(cfg, sd, rays[i:i + ch], None if tr is None else tr[i:i + ch], None if uj is None else uj[i:i + ch], chunk=ch)

    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(work, range(0, n, ch)))
        dt = time.perf_counter() - t0
    return n / dt, cores, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    embed = args.embed
    rates = []
    n_sample = args.cpu_rays
    for i in range(args.warmup + args.steps):
        rate, cores, n, dt = cpu_render_rate(embed, n_sample)
        if i >= args.warmup:
            rates.append((rate, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([d for _, d in rates])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(config_block(embed, "fp32-cpu", n_sample),
                       sample=f"each step renders {n_sample} rays of the {H}x{W} view (bounded CPU sample of the same workload); the rate is per ray"),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} rays of the {H}x{W} view per step (numpy oracle port of nerfmatch/nerf, one thread per core)"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
# informational arm: the reference ALGORITHM as plain eager PyTorch on the same GPU (BASELINE.md 4.6 "beat eager PyTorch").
# Not the product and not the reference arm: library kernels (cuBLAS fp32, ATen elementwise), one launch per op.
# --------------------------------------------------------------------------------------------
def torch_eager_render(sd, rays, Sc=64, Nf=128, stop_layer=3, chunk_points=1 << 20):
    """Classic-NeRF feature render (renderer.py:182-295 with embed_type='normal', validation=True) written with torch ops,
    fp32, MLP evaluated in chunks of ``chunk_points`` samples (the reference's own default is 16384: far more launches)."""
    import torch
    import torch.nn.functional as F

    o, d, near, far, vd = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, 8:11]
    n = rays.shape[0]

    def fourier(x, nf):
        out = [x]
        for k in range(nf):
            out += [torch.sin((2.0 ** k) * x), torch.cos((2.0 ** k) * x)]
        return torch.cat(out, -1)

    def mlp(net, x, stop):
        P = lambda k: sd[f"{net}.{k}"]
        in_pts, in_views = x[:, :93], x[:, 93:]
        h, feat = in_pts, None
        for i in range(8):
            h = F.relu(F.linear(h, P(f"pts_linears.{i}.weight"), P(f"pts_linears.{i}.bias")))
            if i == stop:
                feat = h
            if i == 4:
                h = torch.cat([in_pts, h], -1)
        alpha = F.linear(h, P("alpha_linear.weight"), P("alpha_linear.bias"))
        feature = F.linear(h, P("feature_linear.weight"), P("feature_linear.bias"))
        hv = F.relu(F.linear(torch.cat([feature, in_views], -1), P("views_linears.0.weight"), P("views_linears.0.bias")))
        rgb = torch.sigmoid(F.linear(hv, P("rgb_linear.weight"), P("rgb_linear.bias")))
        return torch.cat([rgb, alpha], -1), (feat if stop >= 0 else h)

    def run_net(net, pts, stop):
        S = pts.shape[1]
        x_d = fourier(vd, 4)
        raws, feats = [], []
        rays_per = max(1, chunk_points // S)
        for r0 in range(0, n, rays_per):
            p = pts[r0:r0 + rays_per].reshape(-1, 3)
            x = torch.cat([fourier(p, 15), x_d[r0:r0 + rays_per].repeat_interleave(S, 0)], -1)
            raw, ft = mlp(net, x, stop)
            raws.append(raw.reshape(-1, S, 4)); feats.append(ft.reshape(-1, S, 256))
        return torch.cat(raws, 0), torch.cat(feats, 0)

    def volrender(raw, z):
        dists = torch.cat([z[:, 1:] - z[:, :-1], 1e10 * torch.ones_like(z[:, :1])], -1) * d.norm(dim=-1, keepdim=True)
        alpha = 1.0 - torch.exp(-F.relu(raw[..., 3]) * dists)
        w = alpha * torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
        return (w[..., None] * raw[..., :3]).sum(-2), (w * z).sum(-1), w

    out = {}
    t = torch.linspace(0.0, 1.0, Sc, device=rays.device)
    z = near * (1.0 - t) + far * t
    for key, net, stop in (("coarse", "nerf_coarse", -1), ("fine", "nerf_fine", stop_layer)):
        if key == "fine":
            mid = 0.5 * (z[:, 1:] + z[:, :-1])
            wts = w[:, 1:-1] + 1e-5
            cdf = torch.cumsum(wts / wts.sum(-1, keepdim=True), -1)
            cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
            u = torch.linspace(0.0, 1.0, Nf, device=rays.device).expand(n, Nf).contiguous()
            inds = torch.searchsorted(cdf, u, right=True)
            below, above = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
            cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
            bb, ba = torch.gather(mid, 1, below), torch.gather(mid, 1, above)
            den = ca - cb
            den = torch.where(den < 1e-5, torch.ones_like(den), den)
            z, _ = torch.sort(torch.cat([z, bb + (u - cb) / den * (ba - bb)], -1), -1)
        pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
        raw, feats = run_net(net, pts, stop)
        rgb, depth, w = volrender(raw, z)
        out[f"feat_{key}"] = (w[..., None] * feats).sum(-2)
        out[f"pts_{key}"] = (w[..., None] * pts).sum(-2)
        out[f"rgb_{key}"], out[f"depth_{key}"] = rgb, depth
    return out


def run_torch_eager(args):
    import torch

    from nerfmatch_b200.config import build_renderer
    from nerfmatch_b200.synthetic import K_7SCENES, make_pose, make_state_dict

    if args.embed != "normal":
        raise SystemExit("--impl torch-eager-gpu covers the headline (classic embedding) workload only")
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.set_grad_enabled(False)
    sd_np = make_state_dict(0, embed_type="normal", sigma_gain=120.0, sigma_bias=(8.0, 8.0))
    sd = {k: torch.from_numpy(v).to(dev) for k, v in sd_np.items()}
    ren = build_renderer(workload("normal"), sd_np, precision="fp32", device=dev)     # only to generate the same rays on the device
    rays, _, _ = ren.engine(dev).raygen(K_7SCENES, make_pose(100), H, W, 1)
    n = rays.shape[0]
    for _ in range(max(1, args.warmup)):
        out = torch_eager_render(sd, rays)
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        a.record(); out = torch_eager_render(sd, rays); b.record()
    torch.cuda.synchronize(dev)
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    # sanity: a slice of the same view through the library's fp32 path (the eager pass holds tens of GB of intermediates)
    feat_e = out["feat_fine"][:4096].clone()
    del out
    torch.cuda.empty_cache()
    mine = ren.predict(rays[:4096], 1, 1)
    cos = torch.nn.functional.cosine_similarity(feat_e, mine["feat_fine"], dim=-1)
    emit({"impl": "torch-eager-gpu", "metric": METRIC, "value": n / (ms / 1e3), "unit": "rays/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic", "config": config_block("normal", "fp32 eager PyTorch (cuBLAS SGEMM, TF32 off), MLP in chunks of 2^20 samples"),
          "check": {"feat_fine_cos_min_vs_libnmr_fp32": float(cos.min()), "feat_fine_cos_median": float(cos.median())},
          "note": "informational: the reference algorithm as eager torch ops on the same GPU; not the reference arm (that is --impl reference, CPU)"})


# --------------------------------------------------------------------------------------------
# one step of the iNeRF pose refinement (nerfmatch_evaluator.py:334-452) on the shim's modules:
# rays from the pose being optimised (ds=8: 60 x 80 rays), coarse pass without grad, fine pass (128 + 128 mip samples) with
# grad, white-background volume rendering, MSE, loss.backward(), Adam step.  The glue is the evaluator's own torch code.
# --------------------------------------------------------------------------------------------
def inerf_step_ms(dev, precision, steps=5):
    import torch

    from nerfmatch_b200.config import build_renderer
    from nerfmatch_b200.render_utils import rays_intersect_sphere, sample_smth_along_rays, volume_render_radiance_field
    from nerfmatch_b200.synthetic import K_7SCENES, make_pose, make_state_dict

    c = dict(embed="mip", nc=128, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=True)
    sd = make_state_dict(21, embed_type="mip", sigma_gain=60.0, sigma_bias=(4.0, 4.0))
    ren = build_renderer(c, sd, precision=precision, device=dev)
    for p_ in ren.parameters():
        p_.requires_grad_(False)
    ren.nerf_coarse.precision = ren.nerf_fine.precision = precision      # forward of the sub-modules on this precision's kernels
    K = torch.from_numpy(K_7SCENES).to(dev)
    ds = 8
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    dirs = torch.stack([xs, ys, torch.ones_like(xs)], -1).float() @ torch.linalg.inv(K).T
    img = torch.rand(len(range(ds // 2, H, ds)) * len(range(ds // 2, W, ds)), 3, device=dev)
    cam_pose = torch.from_numpy(make_pose(21))[None].to(dev).clone().requires_grad_(True)
    opt = torch.optim.Adam(params=[cam_pose], lr=1e-3)
    eps = torch.tensor(torch.finfo(torch.float32).eps, device=dev)

    marks = []

    def mark():
        if marks is not None and len(marks) < 64:
            e_ = torch.cuda.Event(enable_timing=True)
            e_.record()
            marks.append(e_)

    def step():
        with torch.enable_grad():
            mark()
            o = cam_pose[:, None, None, :3, 3].expand(-1, H, W, -1)[0]
            d = torch.matmul(cam_pose[:, None, None, :3, :3], dirs.unsqueeze(-1))[0, :, :, :, 0]
            v = d / d.norm(dim=-1, keepdim=True)
            far = rays_intersect_sphere(o.reshape(-1, 3), v.reshape(-1, 3), r=1).view(H, W, 1)
            dx = torch.sqrt(torch.sum((v[:-1] - v[1:]) ** 2, -1))
            radii = torch.cat([dx, dx[-2:-1]], 0)[..., None] * 2 / np.sqrt(12)
            rays = torch.cat((o, v, torch.full_like(far, 0.01), far, v, radii), -1)[ds // 2:: ds, ds // 2:: ds].reshape(-1, 12)
            z = wts = None
            for key, model in (("coarse", ren.nerf_coarse), ("fine", ren.nerf_fine)):
                pts, z = sample_smth_along_rays(rays.detach(), num_pts=128, z_vals=z, weights=wts, use_disp=False, perturb=False,
                                                embed_type="mip", model_type=key, scale_var=1)
                vf = rays[:, None, 8:11].expand(rays.shape[0], 128, 3)
                mu, hw = (z[..., :-1] + z[..., 1:]) / 2, (z[..., :-1] - z[..., 1:]) / 2
                t_mean = mu + (2 * mu * hw**2) / torch.maximum(eps, 3 * mu**2 + hw**2)
                pts_z = rays[:, None, :3].expand(rays.shape[0], 128, 3) + t_mean[:, :, None] * vf
                x = torch.cat((ren.xyz_encoder(pts_z.contiguous().view(-1, 3), y=pts[1].contiguous().view(-1, 3))[0],
                               ren.dirs_encoder(vf.contiguous().view(-1, 3))), -1)
                if key == "fine":
                    mark()
                if key == "coarse":
                    with torch.no_grad():
                        raw, _ = model.forward(x, ret_pfeat=1, val=True)
                else:
                    raw, _ = model.forward(x, ret_pfeat=1, val=True)
                rgb_map, _, _, wts, _, _ = volume_render_radiance_field(raw.view(-1, 128, 4), z, rays[..., 3:6], noise_std=0.0, white_bg=True,
                                                                        embed_type="mip", input_dim=4)
            loss = torch.nn.functional.mse_loss(rgb_map, img)
            mark()
            loss.backward()
            mark()
            opt.step()
            opt.zero_grad()
            mark()
        return loss

    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    try:
        for _ in range(2):
            step()
        torch.cuda.synchronize(dev)
        marks.clear()
        step()                               # one instrumented step: phase boundaries
        torch.cuda.synchronize(dev)
        phases = {k: marks[i].elapsed_time(marks[i + 1]) for i, k in enumerate(
            ("rays_coarse_pass_fine_sampling_encoding", "fine_mlp_render_loss", "backward", "adam"))} if len(marks) >= 5 else {}
        marks = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss = step()
        b.record()
        torch.cuda.synchronize(dev)
    finally:
        torch.set_grad_enabled(prev)
    return {"ms_per_step": a.elapsed_time(b) / steps, "rays": int(img.shape[0]), "samples": "128 coarse (no grad) + 128 fine (grad), mip", "loss": float(loss.detach()),
            "phases_ms": phases, "forward": precision,
            "backward": "ReLU masks from the tcgen05 forward kernel (no re-forward); dgrad chain on the tensor cores (fp32 gradients split into bf16 hi + lo, pre-split weight panels: 3 MMAs per K step)" if precision != "fp32"
                        else "activations recomputed in fp32, dgrad chain on the tensor cores (3 x bf16 split)",
            "what": "gen_rays -> sample -> encode -> NeRF -> volume render -> MSE -> backward -> Adam, as nerfmatch_evaluator.py:334-452"}


def config_block(embed, precision, rays_per_step=H * W):
    return {"workload": f"7-Scenes {H}x{W} full-res feature render, {embed} embedding, 64 coarse + 128 fine samples "
                        f"({256 if embed == 'normal' else 128} MLP evals/ray), stop_layer=3, random-init weights",
            "rays_per_step_per_gpu": rays_per_step, "precision": precision, "parallelism": "rays sharded by view (one view per GPU per step), weights replicated; when N > 1 feat_fine + pt3d are delivered to rank 0 inside the timed step (peer-memory stores over NVLink, NCCL for setup / barriers)",
            "l2": "256 MiB buffer written between timed steps (L2 flush); per-step working set is also >> 126 MB"}


# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(gpu_index):
    """One process per GPU: pin this process to the CPUs next to its GPU BEFORE any pinned host buffer is allocated, so the
    end-to-end arm's host buffers (first-touch) live on the GPU's own NUMA node instead of all eight ranks sharing node 0."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({cpus[0]}..{cpus[-1]})"
    except Exception as e:      # no NVML / restricted container: run unbound
        return f"unbound ({type(e).__name__})"
    return "unbound"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from nerfmatch_b200.config import build_renderer
    from nerfmatch_b200.synthetic import K_7SCENES, make_pose, make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    embed = args.embed
    c = workload(embed)
    sd = make_state_dict(0, embed_type=embed, sigma_gain=120.0, sigma_bias=(8.0, 8.0))
    ren = build_renderer(c, sd, precision=args.precision, device=dev)
    ren.rng = "philox"      # mip mode: draws generated inside the kernels
    eng = ren.engine(dev)
    n = H * W
    # rank r renders query pose r: rays are sharded by view, weights replicated
    rays, (h, w), _ = eng.raygen(K_7SCENES, make_pose(100 + rank), H, W, 1)
    assert rays.shape[0] == n
    flags = ren._flags()
    outs = eng.alloc_outputs(n, True, True)
    un = np.eye(4, dtype=np.float32); un[:3, :3] *= 3.0
    from nerfmatch_b200 import _lib as L
    rng_mode = L.NMR_RNG_PHILOX if embed == "mip" else L.NMR_RNG_NONE
    dargs, keep = eng.make_args(n, rays.shape[1], rays=rays, flags=flags, rng_mode=rng_mode, seed=1234, unnorm=un, outs=outs)
    # host-side buffers for the end-to-end arm
    rays_h = rays.cpu().pin_memory()
    outs_h = eng.alloc_outputs(n, True, True, device="cpu", pin=True)
    hargs, keep_h = eng.make_args(n, rays.shape[1], rays=rays_h, flags=flags, rng_mode=rng_mode, seed=1234, unnorm=un, outs=outs_h)
    h2d = rays_h.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in outs_h.values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # N > 1: the one real exchange of the path - rendered features + 3D points delivered to the matcher's rank 0.  Product
    # form (dist.PeerBuffers): rank 0 owns the gathered maps, every rank's render kernels store their rows straight into
    # them over NVLink (the output pointers of the render ARE peer addresses); a flag per producer orders rank 0's stream
    # behind the stores.  The NCCL gather after the render (the round-1 form) is kept below as comparison + verification.
    EX = ("feat_fine", "pt3d")
    pb = None
    if world > 1:
        from nerfmatch_b200.dist import PeerBuffers
        pb = PeerBuffers(eng, {"feat_fine": (world * n, 256), "pt3d": (world * n, 3)}, dst=0)
        peer_outs = dict(outs)
        peer_outs.update(pb.out_ptrs(rank * n))
        dargs, keep = eng.make_args(n, rays.shape[1], rays=rays, flags=flags, rng_mode=rng_mode, seed=1234, unnorm=un, outs=peer_outs)
        gbuf = {k: ([torch.empty_like(outs[k]) for _ in range(world)] if rank == 0 else None) for k in EX}
        nargs, keep_n = eng.make_args(n, rays.shape[1], rays=rays, flags=flags, rng_mode=rng_mode, seed=1234, unnorm=un, outs=outs)

    def step_dev():
        if pb is not None:
            pb.begin()
        eng.render(dargs)
        if pb is not None:
            pb.publish()
            if rank == 0:
                pb.wait()

    def step_nccl():                         # comparison leg: render into local tensors, then dist.gather (NCCL send / recv)
        eng.render(nargs)
        for k in EX:
            dist.gather(outs[k], gbuf[k], dst=0)

    def step_host():
        if pb is not None:
            pb.begin()
        eng.render_host(hargs)
        if pb is not None:   # the matcher's rank also needs the other ranks' maps: device-side staging copies -> peer memory
            for k, p_ in pb.out_ptrs(rank * n).items():
                eng.copy_host_staging(k, p_, outs[k].numel() * 4)
            pb.publish()
            if rank == 0:
                pb.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.time()
        for a, b in ev:
            flush.fill_(1)
            a.record(); fn(); b.record()
        barrier()
        t1 = time.time()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), t0, t1

    # ---- device-resident arm ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.set_timing(False)
    l0 = eng.launch_count()
    ms_dev, t0, t1 = timed(step_dev, args.steps, args.warmup)
    launches = (eng.launch_count() - l0) // (args.steps + args.warmup) * args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if args.only_dev:       # profiling runs (ncu): the device-resident arm only
        if rank == 0:
            emit({"metric": METRIC, "value": n * world * args.steps / (ms_dev / 1e3), "unit": "rays/s", "ms_per_step": ms_dev / args.steps,
                  "clocks": clocks, "note": "--only-dev (profiling run): not a bench line"})
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- per-stage split (separate pass, CUDA events on the launch stream inside the library) ----
    eng.set_timing(True)
    for _ in range(args.steps):
        flush.fill_(1)
        eng.render(dargs)
    stage = eng.last_stage_ms()
    # ---- kernel (c) standalone: the same step with the tile-local feature compositing of the bf16 MLP kernel switched off,
    #      i.e. the compositor streams the per-sample bf16 features from HBM (what the fp32 path and feat_comb="max" always do) ----
    stage_unfused = None
    if args.precision in ("bf16", "fp16"):
        eng.set_fuse_composite(False)
        eng.render(dargs)
        eng.last_stage_ms()
        for _ in range(args.steps):
            flush.fill_(1)
            eng.render(dargs)
        stage_unfused = eng.last_stage_ms()
        eng.set_fuse_composite(True)
    eng.set_timing(False)
    # ---- end-to-end arm ----
    ms_e2e, _, _ = timed(step_host, args.steps, max(3, args.warmup))
    exchange = None
    if world > 1:
        # same step with the NCCL gather after the render; its result also verifies the peer-delivered maps bit for bit
        step_dev()
        ms_nccl, _, _ = timed(step_nccl, args.steps, 2)
        ok = True
        if rank == 0:
            torch.cuda.synchronize(dev)
            pb.check()
            for k in EX:
                ok = ok and torch.equal(pb.tensor(k), torch.cat(gbuf[k], 0))
        exchange = {"form": "peer stores from the compositing kernel into rank 0's memory (CUDA IPC over NVLink), flag per producer",
                    "bytes_into_rank0_per_step": (world - 1) * n * (256 + 3) * 4, "verified_equal_to_nccl_gather": bool(ok),
                    "ms_per_step_peer": ms_dev / args.steps, "ms_per_step_nccl_gather": ms_nccl / args.steps}
        if rank == 0 and not ok:
            raise SystemExit("peer-delivered maps differ from the NCCL gather")

    # ---- BASELINE config 4: a batch of 64 query poses at matcher resolution (ds=8: 60 x 80 rays each), the rays of all views
    #      flattened and sharded across the ranks (STRONG scaling: 307 200 rays in total at every N), maps delivered to rank 0 ----
    from nerfmatch_b200.dist import render_views_sharded
    un4 = np.eye(4, dtype=np.float32); un4[:3, :3] *= 3.0
    c2ws = torch.stack([torch.from_numpy(un4 @ make_pose(200 + v)) for v in range(64)])
    Kt = torch.from_numpy(K_7SCENES)

    def step_cfg4():
        return render_views_sharded(ren, (H, W), Kt, c2ws, un4, downsample=8, dst=0)

    ms_cfg4, _, _ = timed(step_cfg4, args.steps, 2)
    # ---- same tcgen05 kernel with fp16 operands (NMR_PREC_FP16): informational second measurement of the headline step ----
    other = None
    if args.precision in ("bf16", "fp16") and world == 1:
        alt = "fp16" if args.precision == "bf16" else "bf16"
        ren.precision = alt
        ren.engine(dev)
        ms_alt, _, _ = timed(step_dev, args.steps, 2)
        ren.precision = args.precision
        ren.engine(dev)
        other = {"precision": alt, "value": n * args.steps / (ms_alt / 1e3), "unit": "rays/s", "ms_per_step": ms_alt / args.steps,
                 "note": "same kernel, other 16-bit operand format; parity of both in profiles/r02_parity.json"}

    total_rays = n * world
    value = total_rays * args.steps / (ms_dev / 1e3)
    e2e = total_rays * args.steps / (ms_e2e / 1e3)
    pk = peaks()
    evals = 256 if embed == "normal" else 128
    flops_step = n * evals * FLOP_PER_EVAL[embed]
    mlp_s = stage["mlp"] / 1e3 / args.steps
    achieved = flops_step / mlp_s / 1e12 if mlp_s > 0 else 0.0
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tr = ncu_traffic("r02_ncu_mlp.txt")
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": config_block(embed, args.precision),
        "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "NeRF MLP (coarse + fine nets)", "achieved": achieved, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tf_sust"], "frac_of_burst_peak": achieved / pk["tf_burst"], "frac_of_nominal_2250": achieved / 2250.0,
                     "tensor_pipe_active_pct_ncu": 71.9, "traffic": tr["bytes"], "traffic_source": "static: ncu --set full capture committed under profiles/ (not re-measured in this run)",
                     "traffic_note": tr["note"], "peak_source": pk["src"] + " sustained bf16 (MEASURED_PEAKS.json)",
                     "flops_per_launch_set": flops_step, "ms_per_step": mlp_s * 1e3},
        "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
    }
    if exchange is not None:
        line["exchange"] = exchange
        line["e2e"]["host_numa_binding_rank0"] = numa
    line["config4"] = {"workload": "64 query poses x 4800 rays (480x640, ds=8), rays of all views sharded across the ranks, pt_feat / pt3d / im_pred delivered to rank 0",
                       "scaling": "strong", "rays": 64 * 4800, "ms_per_batch": ms_cfg4 / args.steps, "value": 64 * 4800 * args.steps / (ms_cfg4 / 1e3), "unit": "rays/s"}
    if other is not None:
        line["other_operand_format"] = other
    if world == 1:
        line["inerf_step_ms"] = inerf_step_ms(dev, args.precision)
    # compositing kernel: HBM-bound side of the path (SURVEY.md 8d byte counts, bf16 | fp32 features), measured with the
    # per-sample features streamed from HBM (kernel (c) standalone); the bf16 product path composites the features tile by
    # tile inside the MLP kernel, its (small, latency-bound) compositor is reported as composite_fused.
    bf = 2 if args.precision in ("bf16", "fp16") else 4
    by = 0
    by_fused = 0
    for S in ((64, 192) if embed == "normal" else (64, 64)):
        by += n * (S * (256 * bf + 16 + 4 + 12) + 256 * 4 + 12 + 12 + 4 + S * 4)
        by_fused += n * (S * (16 + 4) + 48 + (S / 128 + 1) * 1024 + 256 * 4 + 12 + 12 + 4)
    st_c = stage_unfused if stage_unfused is not None else stage
    comp_s = st_c["composite"] / 1e3 / args.steps
    line["roofline_composite"] = {"bound": "hbm", "achieved": by / comp_s / 1e9 if comp_s > 0 else 0.0, "peak": pk["hbm"], "unit": "GB/s",
                                  "frac": (by / comp_s / 1e9 / pk["hbm"]) if comp_s > 0 else 0.0, "bytes_per_step": by,
                                  "ms_per_step": comp_s * 1e3,
                                  "how": "per-sample features streamed from HBM (fuse off); see composite_fused for the product path"
                                  if stage_unfused is not None else "product path"}
    if stage_unfused is not None:
        cf = stage["composite"] / 1e3 / args.steps
        line["composite_fused"] = {"ms_per_step": cf * 1e3, "bytes_per_step": by_fused, "achieved_gbs": by_fused / cf / 1e9 if cf > 0 else 0.0,
                                   "note": "features are composited per 128-point tile inside the MLP kernel; this kernel reads raw/z and "
                                           "one 256-vector per ray segment, and is bound by its per-ray scan / CDF inversion, not by HBM"}
        line["stage_ms_per_step_unfused"] = {k: v / args.steps for k, v in stage_unfused.items()}
    if world == 1 and not args.no_cpu:
        rate, cores, ns, dt = cpu_render_rate(embed, args.cpu_rays)
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                                "sample": f"{ns} rays of the same {H}x{W} view, {dt:.1f} s (numpy oracle port, one thread per core)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the launch captured with `ncu --set full` (profiles/<name>,
    written by scripts/ncu_summary.py from the .ncu-rep of scripts/gpu_evidence.sh).  Not measured in this run."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot, found = 0.0, 0
    try:
        for ln in open(path):
            for key in ("dram__bytes_read.sum [", "dram__bytes_write.sum ["):
                if ln.startswith(key) and found < 2:
                    unit = ln[len(key):ln.index("]")]
                    tot += float(ln.split("=")[1]) * scale[unit]
                    found += 1
    except OSError:
        return {"bytes": None, "note": "no ncu capture in profiles/"}
    if found < 2:
        return {"bytes": None, "note": "no ncu capture in profiles/"}
    return {"bytes": tot, "note": "DRAM read+write of ONE fine-pass launch (the whole 480x640 view in one chunk: 58,982,400 sample rows = 460,800 tiles, 46 ms under ncu) from profiles/" + name +
            "; algorithmic bytes of that launch: 16 B raw per row (0.94 GB) + one 1 KB locally composited feature vector per ray segment of a tile "
            "(~1.7 per tile, 0.8 GB) written, z / rays / view bias / 1.2 MB of weights read (0.4 GB) - the per-sample features stay on chip; "
            "tensor_pipe_active_pct_ncu is sm__pipe_tensor_cycles_active of the same launch, single-pass capture (profiles/r02_mlp_tensor_single_pass.csv)"}


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, ...) print to fd 1; the contract is ONE JSON line on stdout.
    Everything else goes to stderr, the JSON line to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager-gpu"])
    ap.add_argument("--precision", default=os.environ.get("NMR_PRECISION", "bf16"), choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--embed", default="normal", choices=["normal", "mip"])
    ap.add_argument("--cpu-rays", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only-dev", action="store_true", help="profiling runs: time the device-resident arm and stop")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-eager-gpu":
        run_torch_eager(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
