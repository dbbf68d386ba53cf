"""the training-mode forward (perturbed sampling, sigma noise, s_fine / weights_fine;
reference renderer.py:217, 242, 283-286, nerf_trainer.py:140-147) on the GPU against the reference's own run
(tests/golden/train_forward.npz, oracle/gen_golden.py::train_cases) with the draws it consumed injected."""
import ast

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from nmr_helpers import build_renderer, cosine_rows, load_golden, rel_l2  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def _setup(tag, precision):
    g = load_golden("train_forward.npz")
    c = ast.literal_eval(str(g[f"{tag}.case"]))
    sd = make_state_dict(c["seed"], embed_type=c["embed"], sigma_gain=c["sigma_gain"], sigma_bias=c["sigma_bias"])
    ren = build_renderer(c, sd, precision=precision)
    ren.ret_pfeat = tag == "normal"
    return g, ren


@pytest.mark.parametrize("tag", ["mip", "normal"])
def test_training_forward_matches_reference_fp32(tag):
    g, ren = _setup(tag, "fp32")
    draws = tuple(T(g[f"{tag}.{k}"]) for k in ("t_rand", "u", "noise_coarse", "noise_fine"))
    ren._train_draws = lambda n, device: draws                 # the reference's CPU-generator draws (a device generator differs)
    ren.set_training_mode(True)                                # nerf_trainer.py:141-147
    preds = ren.forward(T(g[f"{tag}.rays"]), step=0, ray_id=None)
    assert list(preds.keys()) == list(g[f"{tag}.keys"])
    for k, v in preds.items():
        ref = g[f"{tag}.{k}"]
        assert tuple(v.shape) == ref.shape, k
        if k == "s_fine":
            np.testing.assert_allclose(N(v), ref, rtol=0, atol=2e-5)
        else:
            assert rel_l2(N(v), ref) < 2e-3, (k, rel_l2(N(v), ref))
    # the host-buffer entry takes the same training inputs
    eng = ren.engine(torch.device(DEV))
    n = g[f"{tag}.rays"].shape[0]
    outs = eng.alloc_outputs(n, ren.ret_pfeat, False, device="cpu", pin=True, training=True)
    host = [torch.from_numpy(np.ascontiguousarray(g[f"{tag}.{k}"])).pin_memory() for k in ("rays", "t_rand", "u", "noise_coarse", "noise_fine")]
    args, keep = eng.make_args(n, host[0].shape[1], rays=host[0], t_rand=host[1], u_jitter=host[2], flags=ren._flags(), outs=outs,
                               sigma_noise_coarse=host[3], sigma_noise_fine=host[4])
    eng.render_host(args)
    for k in ("s_fine", "weights_fine", "rgb_fine", "depth_coarse"):
        assert np.array_equal(outs[k].numpy(), N(preds[k])), k


@pytest.mark.parametrize("tag,prec", [("mip", "bf16"), ("normal", "bf16"), ("normal", "fp16")])
def test_training_forward_tensor_core_path(tag, prec):
    g, ren = _setup(tag, prec)
    draws = tuple(T(g[f"{tag}.{k}"]) for k in ("t_rand", "u", "noise_coarse", "noise_fine"))
    ren._train_draws = lambda n, device: draws
    ren.set_training_mode(True)
    preds = ren.forward(T(g[f"{tag}.rays"]))
    assert list(preds.keys()) == list(g[f"{tag}.keys"])
    assert all(torch.isfinite(v).all() for v in preds.values())
    assert np.abs(N(preds["rgb_coarse"]) - g[f"{tag}.rgb_coarse"]).max() < 0.05
    if "feat_coarse" in preds:        # sigma noise switches the in-kernel tile compositing off: features still come out right
        assert cosine_rows(N(preds["feat_coarse"]), g[f"{tag}.feat_coarse"]).min() > 0.999


def test_training_forward_draw_protocol_and_modes():
    """Seeded runs repeat; training vs. eval mode differ; validation=False without training mode is the deterministic
    render plus the s_fine / weights_fine outputs (perturb / noise act only when self.training, renderer.py:217, 242)."""
    g, ren = _setup("normal", "fp32")
    rays = T(g["normal.rays"])
    ren.set_training_mode(True)
    torch.manual_seed(5); a = ren.forward(rays)
    torch.manual_seed(5); b = ren.forward(rays)
    torch.manual_seed(6); c = ren.forward(rays)
    assert all(torch.equal(a[k], b[k]) for k in a) and not torch.equal(a["rgb_fine"], c["rgb_fine"])
    ren.set_training_mode(False)
    d = ren.forward(rays, validation=False)
    e = ren.forward(rays, validation=True)
    assert "s_fine" in d and "pts_fine" not in d and "pts_fine" in e and "s_fine" not in e
    assert torch.equal(d["rgb_fine"], e["rgb_fine"]) and torch.equal(d["depth_coarse"], e["depth_coarse"])
    s = d["s_fine"]
    # s runs from 0 (nearest sample of the batch) to ~1 (farthest; the reference's doubled epsilon overshoots 1 by ~1e-4)
    assert s.min() >= -1e-6 and s.max() <= 1 + 1e-3 and (s[:, 1:] >= s[:, :-1] - 1e-6).all()
    # noise_std = 0 / perturb = False in training mode: no draws are consumed, the result is the deterministic one
    ren.set_training_mode(True)
    ren.noise_std, ren.perturb = 0.0, False
    f = ren.forward(rays)
    assert torch.equal(f["rgb_fine"], d["rgb_fine"])
