"""Every call surface of SURVEY.md 8b that other reference modules reach - the sub-modules the iNeRF refinement calls
directly, the free functions nerfmatch_evaluator / nerf_evaluator import, and the evaluator's construction sequence -
exercised on the GPU against reference fixtures (tests/golden/stages.npz) and the oracle."""
import os
import subprocess
import sys
from argparse import Namespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from nmr_helpers import build_renderer, cosine_rows, load_golden, ref_style_config, rel_l2  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402
from oracle.weights import K_7SCENES, make_pose, make_state_dict, make_unnorm_scene  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def _renderer(embed, app=False, precision="fp32", seed=81, stop_layer=3):
    c = dict(embed=embed, nc=64, nf=128, app=app, stop_layer=stop_layer, feat_comb="lin", white_bg=False)
    sd = make_state_dict(seed, embed_type=embed, app_dim=16 if app else 0, num_frames=7 if app else 0, sigma_gain=60.0, sigma_bias=(4.0, 4.0))
    return build_renderer(c, sd, precision=precision), sd


# ---- encoders as modules (nerfmatch_evaluator.py:377-383) ------------------------------------------------------------------
def test_encoder_modules_match_reference_fixtures():
    from nerfmatch_b200.embedding import FourierEmbedding, PositionalEncodingMIP

    g = load_golden("stages.npz")
    x, var = T(g["enc.x"]), T(g["enc.var"])
    ipe = PositionalEncodingMIP(15)(x, y=var)
    assert isinstance(ipe, tuple) and ipe[1] is None            # the reference returns (enc, enc_var); only [0] is read
    np.testing.assert_allclose(N(ipe[0]), g["enc.ipe"], rtol=0, atol=4e-7)
    np.testing.assert_allclose(N(PositionalEncodingMIP(4)(x)), g["enc.pe_dirs"], rtol=0, atol=4e-7)
    np.testing.assert_allclose(N(FourierEmbedding(15)(x)), g["enc.fourier15"], rtol=0, atol=4e-7)
    np.testing.assert_allclose(N(FourierEmbedding(4)(x)), g["enc.fourier4"], rtol=0, atol=4e-7)
    # leading batch dimensions are preserved, as in the reference
    out = FourierEmbedding(4)(x.reshape(10, 20, 3))
    assert out.shape == (10, 20, 27)
    assert PositionalEncodingMIP(15).get_embedding_dim(3) == 93 and FourierEmbedding(4).get_embedding_dim(3) == 27
    ren, _ = _renderer("mip")
    np.testing.assert_allclose(N(ren.xyz_encoder(x, y=var)[0]), g["enc.ipe"], rtol=0, atol=4e-7)
    np.testing.assert_allclose(N(ren.dirs_encoder(x)), g["enc.pe_dirs"], rtol=0, atol=4e-7)


# ---- NeRF.forward / forward_nerf (nerfmatch_evaluator.py:395-404, renderer.py:119-180) ---------------------------------------
@pytest.mark.parametrize("embed,app", [("mip", False), ("normal", True)])
def test_nerf_module_forward_matches_oracle(embed, app):
    ren, sd = _renderer(embed, app=app)
    cfg = O.RenderConfig(embed_type=embed, app_dim=16 if app else 0, stop_layer=3)
    wts = O.split_state_dict(sd, "nerf_fine.")
    P = 700
    rng = np.random.default_rng(5)
    x = (rng.random((P, cfg.xyz_dim + cfg.dirs_dim + cfg.app_dim), dtype=np.float32) * 2 - 1)
    raw_ref, feat_ref = O.mlp_forward(wts, x, cfg, 3)
    raw, feat = ren.nerf_fine.forward(T(x), ret_pfeat=1, val=True)
    assert raw.shape == (P, 4) and feat.shape == (P, 256)
    assert rel_l2(N(raw), raw_ref) < 1e-5 and rel_l2(N(feat), feat_ref) < 1e-5
    only_raw = ren.nerf_fine.forward(T(x), val=True)             # ret_pfeat=0: a single tensor, like the reference
    assert torch.equal(only_raw, raw)
    # the coarse net taps its last trunk layer (renderer.py:99 passes stop_layer to the fine net only)
    raw_c, feat_c = ren.nerf_coarse.forward(T(x), ret_pfeat=1, val=True)
    raw_cr, feat_cr = O.mlp_forward(O.split_state_dict(sd, "nerf_coarse."), x, cfg, -1)
    assert rel_l2(N(raw_c), raw_cr) < 1e-5 and rel_l2(N(feat_c), feat_cr) < 1e-5
    # the bf16 tensor-core path through the same module surface
    ren.nerf_fine.precision = "bf16"
    ren.nerf_fine._engines.clear()
    raw_b, feat_b = ren.nerf_fine.forward(T(x), ret_pfeat=1, val=True)
    assert cosine_rows(N(feat_b), feat_ref).min() > 0.9995
    assert np.abs(N(raw_b)[:, :3] - raw_ref[:, :3]).max() < 2e-2


@pytest.mark.parametrize("embed,app", [("mip", False), ("normal", True)])
def test_forward_nerf_matches_oracle(embed, app):
    ren, sd = _renderer(embed, app=app)
    cfg = O.RenderConfig(embed_type=embed, app_dim=16 if app else 0, stop_layer=3)
    rays = O.sample_nerf_rays(480, 640, K_7SCENES, make_pose(5), ds=60, embed_type=embed)   # 88 rays
    n, S = rays.shape[0], 64
    viewdirs = rays[:, 8:11]
    app_emb = sd["embedding_a.weight"][np.arange(n) % 7] if app else None
    if embed == "mip":
        _, mean, var = O.sample_gaus_along_rays(rays, S, None)
        pts_arg, pts_np, var_np = (T(mean), T(var)), mean, var
    else:
        pts_np, _ = O.sample_pts_init(rays, S)
        pts_arg, var_np = T(pts_np), None
    x = O.encode_points(cfg, pts_np, var_np, viewdirs, app_emb, S)
    raw_ref, feat_ref = O.mlp_forward(O.split_state_dict(sd, "nerf_fine."), x, cfg, 3)
    out = ren.forward_nerf(ren.nerf_fine, pts_arg, T(viewdirs), ret_pfeat=True, app_emb=None if app_emb is None else T(app_emb), validation=True)
    raw, feats, mask = out
    assert mask is None and raw.shape == (n, S, 4) and feats.shape == (n, S, 256)
    assert rel_l2(N(raw), raw_ref.reshape(n, S, 4)) < 2e-5 and rel_l2(N(feats), feat_ref.reshape(n, S, 256)) < 2e-5
    raw_only = ren.forward_nerf(ren.nerf_fine, pts_arg, T(viewdirs), ret_pfeat=False, app_emb=None if app_emb is None else T(app_emb), validation=True)
    assert torch.equal(raw_only, raw)


# ---- free functions (nerfmatch_evaluator.py:23 imports) --------------------------------------------------------------------------
def test_sample_smth_along_rays_free_function():
    from nerfmatch_b200 import render_utils as RU

    for embed in ("mip", "normal"):
        rays = O.sample_nerf_rays(480, 640, K_7SCENES, make_pose(6), ds=40, embed_type=embed)   # 192 rays
        n = rays.shape[0]
        r = T(rays)
        if embed == "mip":
            torch.manual_seed(3)
            (mean, var), z = RU.sample_smth_along_rays(r, num_pts=128, use_disp=False, perturb=False, embed_type="mip", model_type="coarse", scale_var=1)
            torch.manual_seed(3)
            t_rand = N(torch.rand(n, 129, device=DEV))                                        # the draw of render_utils.py:444
            z_ref, mean_ref, var_ref = O.sample_gaus_along_rays(rays, 128, t_rand)
            np.testing.assert_allclose(N(z), z_ref, rtol=0, atol=2e-7)
            np.testing.assert_allclose(N(mean), mean_ref, rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(N(var), var_ref, rtol=2e-3, atol=1e-12)
            w = np.abs(np.sin(np.arange(n * 128, dtype=np.float32).reshape(n, 128))) ** 6
            torch.manual_seed(4)
            (mean_f, var_f), z_f = RU.sample_smth_along_rays(r, num_pts=128, z_vals=z, weights=T(w), use_disp=False, perturb=False,
                                                             embed_type="mip", model_type="fine", scale_var=1)
            torch.manual_seed(4)
            u_jit = N(torch.empty(n, 129, device=DEV).uniform_(to=(1 / 129 - float(torch.finfo(torch.float32).eps))))
            z_fr, mean_fr, var_fr, _ = O.resample_gaus_along_rays(rays, N(z), w, u_jit)
            np.testing.assert_array_equal(N(z_f), z_fr)
            np.testing.assert_allclose(N(mean_f), mean_fr, rtol=1e-6, atol=1e-7)
            # randomized=False: the deterministic branch (render_utils.py:434-438, 493-496)
            (_, _), z_det = RU.sample_smth_along_rays(r, num_pts=64, embed_type="mip", model_type="coarse", randomized=False)
            np.testing.assert_allclose(N(z_det), O.sample_gaus_along_rays(rays, 64, None)[0], rtol=0, atol=2e-7)
        else:
            pts, z = RU.sample_smth_along_rays(r, num_pts=64, use_disp=False, perturb=False, embed_type="normal", model_type="coarse")
            pts_ref, z_ref = O.sample_pts_init(rays, 64)
            np.testing.assert_array_equal(N(z), z_ref)
            np.testing.assert_allclose(N(pts), pts_ref, rtol=1e-6, atol=1e-7)
            w = np.abs(np.cos(np.arange(n * 64, dtype=np.float32).reshape(n, 64))) ** 8
            pts_f, z_f = RU.sample_smth_along_rays(r, num_pts=128, z_vals=z, weights=T(w), use_disp=False, perturb=False, embed_type="normal", model_type="fine")
            pts_fr, z_fr, _ = O.sample_pts_fine(rays, z_ref, w, 128)
            np.testing.assert_array_equal(N(z_f), z_fr)
            np.testing.assert_allclose(N(pts_f), pts_fr, rtol=1e-6, atol=1e-7)
            # use_disp=True is the reference's DEFAULT for this function
            _, z_d = RU.sample_smth_along_rays(r, num_pts=64, embed_type="normal", model_type="coarse")
            np.testing.assert_allclose(N(z_d), O.sample_pts_init(rays, 64, use_disp=True)[1], rtol=2e-6, atol=0)


def test_volume_render_radiance_field_free_function():
    from nerfmatch_b200 import render_utils as RU

    g = load_golden("stages.npz")
    raw = T(g["volrender.raw"])
    d = T(g["mip_resample.d"])
    for mode, zz in (("mip", g["mip_resample.t"]), ("normal", g["normal_resample.z"])):
        for wb in (False, True):
            rgb, disp, acc, weights, depth, last = RU.volume_render_radiance_field(raw, T(zz), d, noise_std=0.0, white_bg=wb, embed_type=mode, input_dim=4)
            tag = f"volrender.{mode}.{'white' if wb else 'black'}"
            assert last is None
            np.testing.assert_allclose(N(weights), g[tag + ".weights"], rtol=2e-5, atol=1e-7)
            np.testing.assert_allclose(N(rgb), g[tag + ".rgb"], rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(N(acc), g[tag + ".acc"], rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(N(depth), g[tag + ".depth"], rtol=2e-5, atol=2e-6)
            ok = np.isfinite(g[tag + ".disp"])
            np.testing.assert_allclose(N(disp)[ok], g[tag + ".disp"][ok], rtol=1e-4)


# ---- construction sequence of the evaluators (nerf_evaluator.py:119-146, 174-176) --------------------------------------------
def test_evaluator_construction_sequence():
    """hyper-parameters -> ctor -> .to(device) -> .eval() -> strip 'model.' -> load_state_dict(strict=False) -> .unnorm_scene ->
    predict(), exactly as load_nerf_render_from_ckpt does, against the oracle."""
    from nerfmatch_b200 import NerfRenderer

    hp = vars(ref_style_config("mip"))                              # ckpt["hyper_parameters"]: a dict of nested Namespaces
    sd = make_state_dict(83, embed_type="mip", sigma_gain=60.0, sigma_bias=(4.0, 4.0))
    ckpt_sd = {"model." + k: torch.from_numpy(v) for k, v in sd.items()}
    ckpt_sd["loss_fn.some_buffer"] = torch.zeros(3)                 # non-model entries are dropped by the prefix filter
    vocab_num = ckpt_sd["model.embedding_a.weight"].shape[0] if "model.embedding_a.weight" in ckpt_sd else -1
    config = Namespace(**hp)
    render = NerfRenderer(config, num_frames=vocab_num, training=False, stop_layer=3)
    render.to(DEV)
    render.eval()
    new_sd = {k.replace("model.", ""): v for k, v in ckpt_sd.items() if k.startswith("model.")}
    new_sd["nerf_fine.extra_head.weight"] = torch.zeros(2, 2)       # strict=False tolerates unexpected keys
    state = render.load_state_dict(new_sd, strict=False)
    assert state.missing_keys == [] and state.unexpected_keys == ["nerf_fine.extra_head.weight"]
    render.unnorm_scene = torch.from_numpy(make_unnorm_scene())
    render.ret_pfeat, render.feat_comb = True, "lin"                # poked from outside (nerf_evaluator.py:310-311)
    render.precision = "fp32"
    rays = O.sample_nerf_rays(480, 640, K_7SCENES, make_pose(8), ds=60, embed_type="mip")
    n = rays.shape[0]
    torch.manual_seed(11)
    preds = render.predict(T(rays), 11, 8)
    torch.manual_seed(11)
    t_rand = N(torch.rand(n, 65, device=DEV))
    u_jit = N(torch.empty(n, 65, device=DEV).uniform_(to=(1 / 65 - float(torch.finfo(torch.float32).eps))))
    ref = O.render_rays(O.RenderConfig(embed_type="mip", stop_layer=3), sd, rays, t_rand, u_jit)
    assert list(preds) == ["feat_coarse", "pts_coarse", "rgb_coarse", "depth_coarse", "feat_fine", "pts_fine", "rgb_fine", "depth_fine"]
    assert preds["rgb_fine"].shape == (8, 11, 3) and preds["depth_fine"].shape == (8, 11, 1)
    for k in ("feat_fine", "pts_fine"):
        assert rel_l2(N(preds[k]), ref[k]) < 1e-3, k
    # attributes the evaluators read / poke keep their reference meaning
    assert render.mip_var_scale == -1 and render.embed_type == "mip" and render.training is False
    render.mip_var_scale = 2.0                                       # nerf_evaluator.py:195-196: picked up on the next call
    torch.manual_seed(11)
    p2 = render.predict(T(rays), 11, 8)
    ref2 = O.render_rays(O.RenderConfig(embed_type="mip", stop_layer=3, mip_var_scale=2.0), sd, rays, t_rand, u_jit)
    assert rel_l2(N(p2["feat_fine"]), ref2["feat_fine"]) < 1e-3 and not torch.equal(p2["feat_fine"], preds["feat_fine"])


def test_match_small_problem_sizes():
    """n < 255: threads of the row-statistics block that see no column must not poison the softmax merge (NaN)."""
    from nerfmatch_b200.matching import coarse_match
    from oracle import match_oracle as MO

    rng = np.random.default_rng(3)
    for m, n in ((40, 100), (3, 7), (300, 254), (1, 1)):
        pt = rng.standard_normal((n, 64)).astype(np.float32)
        im = (pt[rng.integers(0, n, m)] + 0.3 * rng.standard_normal((m, 64))).astype(np.float32)
        out = coarse_match(T(im), T(pt), temperature=10.0, temp_type="mul", mutual=True, match_thres=0.0)
        conf = N(out["conf_matrix"][0])
        assert np.isfinite(conf).all()
        ref_conf = MO.coarse_matching(im, pt, 10.0, "mul")
        np.testing.assert_allclose(conf, ref_conf, rtol=2e-4, atol=1e-7)
        i_ref, j_ref, _ = MO.extract_mutual_matches(ref_conf, mutual=True, threshold=0.0)
        assert out["pred_num"] == len(i_ref) and out["pred_num"] > 0
        np.testing.assert_array_equal(N(out["match_ids"][1]), i_ref)
        np.testing.assert_array_equal(N(out["match_ids"][2]), j_ref)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_render_views_sharded_two_gpus():
    """BASELINE config 4 path on two ranks (NCCL): the gathered maps equal the unsharded render of every view."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tests", "dist_worker_gpu.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SHARDED_OK" in res.stdout


def test_single_model_and_pfeat_mask_options():
    """config.render.single_model (the fine net renders both passes with num_pts_fine samples, renderer.py:76-87, 191-201) and
    NerfRenderer.pfeat_mask (features for the masked rays only, nerf_trainer.py:45-47) against the oracle."""
    from nerfmatch_b200 import NerfRenderer

    for embed, nf in (("mip", 64), ("normal", 96)):
        cfg = ref_style_config(embed, nc=64, nf=nf)
        cfg.render.single_model = True
        sd = make_state_dict(85, embed_type=embed, sigma_gain=60.0, sigma_bias=(4.0, 4.0))
        ren = NerfRenderer(cfg, training=False, stop_layer=-1, precision="fp32")
        ren.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items() if not k.startswith("nerf_coarse")}, strict=True)
        ren = ren.to(DEV).eval()
        ren.ret_pfeat, ren.rng = True, "none"
        rays = O.sample_nerf_rays(480, 640, K_7SCENES, make_pose(9), ds=40, embed_type=embed)   # 192 rays
        sd_single = dict(sd)
        for k in list(sd):
            if k.startswith("nerf_fine."):
                sd_single["nerf_coarse." + k[len("nerf_fine."):]] = sd[k]
        ocfg = O.RenderConfig(embed_type=embed, num_pts_coarse=nf, num_pts_fine=nf, stop_layer=-1, stop_layer_coarse=-1)
        ref = O.render_rays(ocfg, sd_single, rays)
        out = ren.predict(T(rays), 1, 1)
        for k in ("feat_coarse", "feat_fine", "pts_fine", "depth_fine"):
            assert rel_l2(N(out[k]), ref[k]) < 1e-3, (embed, k, rel_l2(N(out[k]), ref[k]))
        mask = torch.zeros(12, 16, 1, dtype=torch.bool)
        mask[1::4, 2::4] = True                                     # init_pfeat_mask-style strided selection
        ren.pfeat_mask = mask
        sub = ren.forward(T(rays), validation=True)
        sel = mask.flatten().numpy()
        assert sub["feat_fine"].shape == (int(sel.sum()), 256) and sub["pts_fine"].shape == (192, 3)
        assert torch.equal(sub["feat_fine"], out["feat_fine"][T(sel)]) and torch.equal(sub["pts_fine"], out["pts_fine"])
        ren.pfeat_mask = None
