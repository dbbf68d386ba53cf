"""the differentiable path of the iNeRF pose refinement on the GPU.

The loop of nerfmatch/nerfmatch_evaluator.py:288-500 (ray generation from the pose being optimised, encodings, fine NeRF,
white-background volume rendering, MSE, loss.backward(), Adam) is replayed with the shim's sub-modules and free functions
under torch autograd - the glue lines below restate the evaluator's own torch code - and checked against
tests/golden/inerf_grad.npz, which oracle/gen_golden.py::inerf_cases produced with torch autograd through the REFERENCE's
functions.  Tolerances are those of the oracle's own test (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from nmr_helpers import build_renderer, load_golden, rel_l2  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def gen_rays(poses, width, height, z_near, K, ds):
    """nerfmatch_evaluator.py:232-285 (gen_rays), device-resident."""
    from nerfmatch_b200.render_utils import rays_intersect_sphere

    device = poses.device
    ys, xs = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
    xys = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).float()
    dirs = xys @ torch.linalg.inv(K.squeeze()).T
    cam_centers = poses[:, None, None, :3, 3].expand(-1, height, width, -1)[0]
    cam_raydir = torch.matmul(poses[:, None, None, :3, :3], dirs.unsqueeze(-1))[0, :, :, :, 0]
    viewdirs = cam_raydir / cam_raydir.norm(dim=-1, keepdim=True)
    cam_nears = torch.tensor(z_near, device=device).view(1, 1, 1).expand(height, width, -1)
    z_far = rays_intersect_sphere(cam_centers.reshape(-1, 3), viewdirs.reshape(-1, 3), r=1)
    cam_fars = z_far.view(height, width, -1)
    dx = torch.sqrt(torch.sum((viewdirs[:-1, :, :] - viewdirs[1:, :, :]) ** 2, -1))
    dx = torch.cat([dx, dx[-2:-1, :]], 0)
    radii = dx[..., None] * 2 / np.sqrt(12)
    rays = torch.cat((cam_centers, viewdirs, cam_nears, cam_fars, viewdirs, radii), dim=-1)
    return rays[ds // 2:: ds, ds // 2:: ds].reshape((-1, rays.shape[-1]))


def fine_pass(renderer, rays, z_vals, var):
    """nerfmatch_evaluator.py:345-420 for the fine model, with the (constant) samples given."""
    from nerfmatch_b200.render_utils import volume_render_radiance_field

    rays_d, viewdirs = rays[..., 3:6], rays[..., 8:11]
    S = z_vals.shape[-1] - 1
    viewdirs_flat = viewdirs[..., None, :].expand(rays.shape[0], S, 3)
    mu = (z_vals[..., :-1] + z_vals[..., 1:]) / 2
    hw = (z_vals[..., :-1] - z_vals[..., 1:]) / 2
    eps = torch.tensor(torch.finfo(torch.float32).eps, device=rays.device)
    t_mean = mu + (2 * mu * hw**2) / torch.maximum(eps, 3 * mu**2 + hw**2)
    var_flat = var.contiguous().view(-1, 3)
    pts_z = rays[:, None, :3].expand(rays.shape[0], S, 3) + (t_mean[:, :, None] * viewdirs_flat)
    inputs = renderer.xyz_encoder(pts_z.contiguous().view(-1, 3), y=var_flat)[0]
    inputs = torch.cat((inputs, renderer.dirs_encoder(viewdirs_flat.contiguous().view(-1, 3))), dim=-1)
    if inputs.requires_grad:
        inputs.retain_grad()
    raw_outs, feats = renderer.nerf_fine.forward(inputs, ret_pfeat=1, val=True)
    if raw_outs.requires_grad:
        raw_outs.retain_grad()
    rendered = volume_render_radiance_field(raw_outs.view(-1, S, raw_outs.shape[-1])[..., :], z_vals, rays_d, noise_std=0.0, white_bg=True,
                                            embed_type="mip", input_dim=4)
    return rendered, inputs, raw_outs, feats


def _setup(tag):
    g = load_golden("inerf_grad.npz")
    seed, gain, bias = g[f"{tag}.seed_gain_bias"]
    sd = make_state_dict(int(seed), embed_type="mip", sigma_gain=float(gain), sigma_bias=(float(bias), float(bias)))
    c = dict(embed="mip", nc=128, nf=128, app=False, stop_layer=3, feat_comb="lin", white_bg=True)
    ren = build_renderer(c, sd, precision="fp32")
    for p in ren.parameters():
        p.requires_grad_(False)
    return g, sd, ren


@pytest.mark.parametrize("tag,prec", [("smooth", "fp32"), ("peaky", "fp32"), ("smooth", "fp16"), ("smooth", "bf16")])
def test_inerf_fine_pass_gradient_matches_reference_autograd(tag, prec):
    """prec = fp32: the parity path (activations recomputed in the backward).  fp16 / bf16: the sub-module's forward runs on the
    tcgen05 kernel, which emits the ReLU masks, and the backward is the dgrad chain alone - same gradient up to the operand
    rounding of the forward (the masks and rgb it differentiates belong to the 16-bit forward)."""
    g, sd, ren = _setup(tag)
    if prec != "fp32":
        ren.nerf_fine.precision = prec
    tol = {"fp32": 1.0, "fp16": 8.0, "bf16": 60.0}[prec]
    H, W, ds = [int(v) for v in g[f"{tag}.hwds"]]
    K = T(g[f"{tag}.K"])
    img = T(g[f"{tag}.img"])
    torch.set_grad_enabled(False)                      # the evaluators disable grad globally (nerf_evaluator.py:155) ...
    try:
        cam_pose = T(g[f"{tag}.pose"])[None].clone()
        cam_pose.requires_grad = True
        with torch.enable_grad():                      # ... and the refinement re-enables it locally (nerfmatch_evaluator.py:337)
            rays = gen_rays(cam_pose, W, H, 0.01, K, ds)
            np.testing.assert_allclose(N(rays), g[f"{tag}.rays"], rtol=2e-5, atol=2e-6)
            (rgb_map, disp, acc, weights, depth, last), inputs, raw_outs, feats = fine_pass(ren, rays, T(g[f"{tag}.z_fine"]), T(g[f"{tag}.var_fine"]))
            loss = torch.nn.MSELoss(reduction="mean")(rgb_map, img)
            loss.backward()
    finally:
        torch.set_grad_enabled(True)
    assert abs(loss.item() - float(g[f"{tag}.loss"])) < 2e-5 * tol * max(1.0, abs(loss.item()))
    np.testing.assert_allclose(N(rgb_map), g[f"{tag}.rgb_map"], rtol=0, atol=5e-5 * tol)
    assert rel_l2(N(raw_outs.grad), g[f"{tag}.grad_raw"]) < 5e-4 * tol
    assert rel_l2(N(inputs.grad)[:1024], g[f"{tag}.grad_inputs_head"]) < 3e-3 * (1.0 if prec == "fp32" else (3.0 if prec == "fp16" else 15.0))
    ref = g[f"{tag}.grad_pose"]
    got = N(cam_pose.grad[0])
    assert np.abs(got[3]).max() == 0
    assert np.abs(got[:3] - ref[:3]).max() < (1e-2 if prec == "fp32" else (2e-2 if prec == "fp16" else 8e-2)) * np.abs(ref).max(), (got[:3], ref[:3])
    if prec != "fp32":
        return
    # one Adam step (nerfmatch_evaluator.py:334, 450-452): the reference's pose update from the reference's gradient
    cam2 = T(g[f"{tag}.pose"])[None].clone().requires_grad_(True)
    opt = torch.optim.Adam(params=[cam2], lr=0.001)
    cam2.grad = cam_pose.grad.clone()
    opt.step()
    expect = g[f"{tag}.pose"] - 0.001 * ref / (np.abs(ref) + 1e-8)
    big = np.abs(ref) > 1e-2 * np.abs(ref).max()       # entries whose sign is determined beyond the comparison tolerance
    np.testing.assert_allclose(N(cam2[0])[big], expect[big], rtol=0, atol=2e-5)


def test_stage_backward_entry_points_vs_finite_differences():
    """nmr_stage_*_backward against central differences of the forward stages (fp32 forward, float64 accumulation of the
    probe): encodings in all three modes, the MLP (incl. the tapped-feature gradient and a skip-layer tap), compositing with
    every incoming gradient and both sampling modes."""
    from nerfmatch_b200.engine import Engine, EngineConfig

    rng = np.random.default_rng(0)
    # ---- encodings ----
    for embed, kind, use_var in (("mip", 0, True), ("mip", 1, False), ("normal", 0, False), ("normal", 1, False)):
        eng = Engine(EngineConfig(embed_type=embed, xyz_num_freqs=6, dirs_num_freqs=4), torch.device(DEV))   # low octaves: FD-friendly
        P = 64
        x = (rng.random((P, 3)) * 2 - 1).astype(np.float32)
        var = (rng.random((P, 3)) * 1e-3).astype(np.float32) if use_var else None
        out = eng.stage_encode(kind, T(x), None if var is None else T(var))
        gout = rng.standard_normal(out.shape).astype(np.float32)
        gx = N(eng.stage_encode_backward(kind, T(x), None if var is None else T(var), T(gout)))
        h = 1e-3
        fd = np.zeros_like(x, np.float64)
        for c in range(3):
            xp, xm = x.copy(), x.copy()
            xp[:, c] += h
            xm[:, c] -= h
            d = (N(eng.stage_encode(kind, T(xp), None if var is None else T(var))).astype(np.float64)
                 - N(eng.stage_encode(kind, T(xm), None if var is None else T(var))).astype(np.float64)) / (xp[:, c] - xm[:, c]).astype(np.float64)[:, None]
            fd[:, c] = (d * gout).sum(-1)
        assert rel_l2(gx, fd) < 5e-3, (embed, kind, rel_l2(gx, fd))
    # ---- MLP: torch autograd (float64, CPU) through the same network written with F.linear (models/nerf.py:94-144) ----
    import torch.nn.functional as F

    import os

    for embed, app, stop, tc in (("mip", 0, 3, "1"), ("normal", 16, 4, "1"), ("mip", 0, -1, "0"), ("normal", 0, 7, "1"), ("normal", 16, 4, "0")):
        # NMR_GEMM_TC (read at nmr_create): 1 = dgrad chain on the tensor cores (fp32 operands split into bf16 hi + lo, three
        # MMAs per K step: ~2^-16 per product), 0 = fp32 SIMT GEMMs
        os.environ["NMR_GEMM_TC"] = tc
        cfg = O.RenderConfig(embed_type=embed, app_dim=app)
        sd = make_state_dict(5, embed_type=embed, app_dim=app, num_frames=3 if app else 0, sigma_gain=40.0, sigma_bias=2.0)
        wts = O.split_state_dict(sd, "nerf_fine.")
        eng = Engine(EngineConfig(embed_type=embed, app_dim=app, precision="fp32"), torch.device(DEV))
        os.environ.pop("NMR_GEMM_TC")
        eng.set_weights(0, {k: torch.from_numpy(v) for k, v in wts.items()})
        P, D = 300, cfg.xyz_dim + cfg.dirs_dim + app
        x = (rng.random((P, D)) * 2 - 1).astype(np.float32)
        g_raw = rng.standard_normal((P, 4)).astype(np.float32)
        g_feat = rng.standard_normal((P, 256)).astype(np.float32) * 0.1
        gx = N(eng.stage_mlp_backward(0, T(x), stop, T(g_raw), T(g_feat))).astype(np.float64)
        W = {k: torch.from_numpy(v).double() for k, v in wts.items()}
        xt = torch.from_numpy(x).double().requires_grad_(True)
        in_pts, in_views = xt[:, :cfg.xyz_dim], xt[:, cfg.xyz_dim:]
        h, feat = in_pts, None
        for i in range(8):
            h = F.relu(F.linear(h, W[f"pts_linears.{i}.weight"], W[f"pts_linears.{i}.bias"]))
            if i == stop:
                feat = h
            if i == 4:
                h = torch.cat([in_pts, h], -1)
        feat = h if stop < 0 else feat
        alpha = F.linear(h, W["alpha_linear.weight"], W["alpha_linear.bias"])
        hv = F.relu(F.linear(torch.cat([F.linear(h, W["feature_linear.weight"], W["feature_linear.bias"]), in_views], -1),
                             W["views_linears.0.weight"], W["views_linears.0.bias"]))
        raw = torch.cat([torch.sigmoid(F.linear(hv, W["rgb_linear.weight"], W["rgb_linear.bias"])), alpha], -1)
        ((raw * torch.from_numpy(g_raw).double()).sum() + (feat * torch.from_numpy(g_feat).double()).sum()).backward()
        ref = xt.grad.numpy()
        # a unit whose pre-activation is within fp32 round-off of zero can take the other ReLU branch: per-row comparison
        err = np.linalg.norm(gx - ref, axis=1) / np.linalg.norm(ref, axis=1)
        assert np.median(err) < (2e-4 if tc == "1" else 1e-5) and (err < 2e-3).mean() > 0.98, (embed, stop, tc, np.median(err), (err < 2e-3).mean())
    # ---- compositing ----
    for embed in ("mip", "normal"):
        eng = Engine(EngineConfig(embed_type=embed), torch.device(DEV))
        n, S = 40, 48
        raw = rng.random((n, S, 4)).astype(np.float32)
        raw[..., 3] = (raw[..., 3] - 0.3) * 30
        nz = S + 1 if embed == "mip" else S
        z = np.sort(rng.random((n, nz)).astype(np.float32) * 1.2 + 0.05, -1)
        d = rng.standard_normal((n, 3)).astype(np.float32)
        g_rgb, g_dep, g_acc = rng.standard_normal((n, 3)).astype(np.float32), rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
        g_w = rng.standard_normal((n, S)).astype(np.float32)
        for flags in (0, 4):
            g_raw_t, g_d_t = eng.stage_composite_backward(T(raw), T(z), T(d), flags=flags, grad_rgb=T(g_rgb), grad_depth=T(g_dep), grad_acc=T(g_acc), grad_weights=T(g_w))

            def loss_of(raw_, d_):
                o = eng.stage_composite(T(raw_), T(z), T(d_), flags=flags)
                return float((N(o["rgb"]).astype(np.float64) * g_rgb).sum() + (N(o["depth"]).astype(np.float64) * g_dep).sum()
                             + (N(o["acc"]).astype(np.float64) * g_acc).sum() + (N(o["weights"]).astype(np.float64) * g_w).sum())

            for name, base, grad in (("raw", raw, N(g_raw_t)), ("d", d, N(g_d_t))):
                v = rng.standard_normal(base.shape).astype(np.float32)
                if name == "raw" and embed == "normal":
                    v[:, -1, 3] = 0        # the classic last interval is 1e10 wide: d alpha / d sigma is a delta function there
                h = 1e-3
                fd = (loss_of(raw + h * v, d) - loss_of(raw - h * v, d)) / (2 * h) if name == "raw" else (loss_of(raw, d + h * v) - loss_of(raw, d - h * v)) / (2 * h)
                an = float((grad.astype(np.float64) * v).sum())
                assert abs(an - fd) < 2e-2 * max(1.0, abs(fd)), (embed, flags, name, an, fd)


def test_inerf_refinement_loop_runs():
    """The whole refinement loop shape (coarse pass without grad, fine pass with grad, Adam on the pose; evaluator lines
    :334-452) on the shim's modules: six steps run, every step yields a finite loss and a non-zero pose gradient, and Adam
    moves each pose entry by at most the learning rate per step.  (A random-init field renders a nearly view-independent
    image - the loss is ~1e-8 everywhere - so convergence itself is not asserted; the gradient's value is pinned by the
    reference-autograd test above.)"""
    from nerfmatch_b200.render_utils import sample_smth_along_rays, volume_render_radiance_field

    g, sd, ren = _setup("smooth")
    H, W, ds = [int(v) for v in g["smooth.hwds"]]
    K = T(g["smooth.K"])
    pose_gt = T(g["smooth.pose"])[None]
    torch.set_grad_enabled(False)
    try:
        def render(cam_pose, seed):
            rays = gen_rays(cam_pose, W, H, 0.01, K, ds)
            rays_d = rays[..., 3:6]
            z_vals = weights = None
            torch.manual_seed(seed)
            for key, model in (("coarse", ren.nerf_coarse), ("fine", ren.nerf_fine)):
                pts, z_vals = sample_smth_along_rays(rays.detach(), num_pts=128, z_vals=z_vals, weights=weights, use_disp=False, perturb=False,
                                                     embed_type="mip", model_type=key, scale_var=1)
                if key == "coarse":
                    with torch.no_grad():
                        (rgb_map, _, _, weights, _, _), _, _, _ = fine_pass_with(ren, model, rays, z_vals, pts[1])
                else:
                    (rgb_map, _, _, weights, _, _), _, _, _ = fine_pass_with(ren, model, rays, z_vals, pts[1])
            return rgb_map

        def fine_pass_with(renderer, model, rays, z_vals, var):
            keep = renderer.nerf_fine
            try:
                renderer.nerf_fine = model
                return fine_pass(renderer, rays, z_vals, var)
            finally:
                renderer.nerf_fine = keep

        target = render(pose_gt, 0).detach()
        cam_pose = pose_gt.clone()
        cam_pose[0, :3, 3] += torch.tensor([0.004, -0.003, 0.005], device=DEV)
        cam_pose.requires_grad = True
        opt = torch.optim.Adam(params=[cam_pose], lr=0.001)
        losses, gnorm, moved = [], [], []
        for j in range(6):
            before = cam_pose.detach().clone()
            with torch.enable_grad():
                rgb_map = render(cam_pose, j)
                loss = torch.nn.MSELoss(reduction="mean")(rgb_map, target)
                loss.backward()
                gnorm.append(cam_pose.grad.abs().max().item())
                assert cam_pose.grad[0, 3].abs().max().item() == 0
                opt.step()
                opt.zero_grad()
            losses.append(loss.item())
            moved.append((cam_pose.detach() - before).abs().max().item())
    finally:
        torch.set_grad_enabled(True)
    assert np.isfinite(losses).all() and min(gnorm) > 0
    assert max(moved) <= 4e-3 and min(moved) > 0, moved      # Adam: |step| <= lr (1 - b1) / sqrt(1 - b2) = 3.2 lr, = lr on step one
