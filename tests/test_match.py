"""NeRFMatch-Mini coarse matching: oracle vs reference fixtures (CPU), CUDA vs oracle (GPU)."""
import numpy as np
import pytest

from nmr_helpers import load_golden
from oracle import match_oracle as M

CASES = [("mul10", 10.0, "mul", False), ("div01_masked", 0.1, "div", True)]


@pytest.mark.parametrize("tag,temp,ttype,masks", CASES)
def test_match_oracle_matches_reference(tag, temp, ttype, masks):
    g = load_golden("match.npz")
    conf = M.coarse_matching(g["im"], g["pt"], temp, ttype, g["im_mask"] if masks else None, g["pt_mask"] if masks else None)
    np.testing.assert_allclose(conf, g[f"{tag}.conf"], rtol=2e-4, atol=1e-9)
    for mutual in (False, True):
        for thr in (0.0, 0.2):
            k = f"{tag}.{'mutual' if mutual else 'rowmax'}.thr{thr}"
            i, j, c = M.extract_mutual_matches(g[f"{tag}.conf"], mutual, thr)     # on the reference's conf: exact
            assert np.array_equal(i, g[k + ".i"]) and np.array_equal(j, g[k + ".j"])
            np.testing.assert_array_equal(c, g[k + ".mconf"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,temp,ttype,masks", CASES)
def test_cuda_coarse_match_vs_reference(tag, temp, ttype, masks):
    torch = pytest.importorskip("torch")
    from nerfmatch_b200.matching import coarse_match

    g = load_golden("match.npz")
    im, pt = torch.from_numpy(g["im"]).cuda(), torch.from_numpy(g["pt"]).cuda()
    imk = torch.from_numpy(g["im_mask"]).cuda() if masks else None
    ptk = torch.from_numpy(g["pt_mask"]).cuda() if masks else None
    for mutual in (False, True):
        for thr in (0.0, 0.2):
            out = coarse_match(im[None], pt[None], temp, ttype, imk, ptk, mutual=mutual, match_thres=thr)
            conf = out["conf_matrix"][0].cpu().numpy()
            np.testing.assert_allclose(conf, g[f"{tag}.conf"], rtol=3e-4, atol=1e-9)
            k = f"{tag}.{'mutual' if mutual else 'rowmax'}.thr{thr}"
            i, j = out["match_ids"][1].cpu().numpy(), out["match_ids"][2].cpu().numpy()
            # matches are decided by exact float ties on OUR conf: identical to the oracle's decision on that conf ...
            io, jo, co = M.extract_mutual_matches(conf, mutual, thr)
            assert np.array_equal(i, io) and np.array_equal(j, jo)
            np.testing.assert_array_equal(out["mconf"].cpu().numpy(), co)
            # ... and to the reference's matches except where a confidence sits within round-off of the threshold
            ref = set(zip(g[k + ".i"].tolist(), g[k + ".j"].tolist()))
            got = set(zip(i.tolist(), j.tolist()))
            assert len(ref ^ got) <= max(2, len(ref) // 100), (k, len(ref ^ got), len(ref))


@pytest.mark.gpu
def test_cuda_coarse_match_full_size_properties():
    """BASELINE config 3 shape: 60x107 image cells vs 6420 rendered points, d = 256."""
    torch = pytest.importorskip("torch")
    from nerfmatch_b200.matching import coarse_match

    gen = torch.Generator("cuda").manual_seed(3)
    n = m = 60 * 107
    pt = torch.randn(n, 256, device="cuda", generator=gen)
    perm = torch.randperm(n, device="cuda", generator=gen)
    im = pt[perm] + 0.3 * torch.randn(m, 256, device="cuda", generator=gen)
    out = coarse_match(im, pt, 10.0, "mul", mutual=True, match_thres=0.2)
    conf = out["conf_matrix"][0]
    assert (conf >= 0).all() and (conf <= 1).all()
    assert (conf.sum(0) <= 1 + 1e-4).all() and (conf.sum(1) <= 1 + 1e-4).all()    # each softmax factor sums to 1
    i, j = out["match_ids"][1], out["match_ids"][2]
    assert out["pred_num"] > 0.9 * m                                                 # noisy copies are recovered ...
    assert (perm[i] == j).float().mean() > 0.999                                     # ... at the right place
    assert len(torch.unique(j)) == len(j) and (i[1:] > i[:-1]).all()                 # mutual -> one-to-one, ordered


@pytest.mark.gpu
def test_similarity_gemm_tensor_core_vs_simt():
    """The matcher's one contraction runs on the tensor cores (fp32 operands split into bf16 hi + lo, 3 MMAs per K step);
    NMR_GEMM_TC=0 selects the fp32 SIMT GEMM.  Same matches, confidences equal to ~1e-4 relative (exp(10 * 2^-16))."""
    import os

    import torch

    import nerfmatch_b200.matching as MM

    rng = np.random.default_rng(11)
    m, n, d = 1000, 1300, 256
    pt = rng.standard_normal((n, d)).astype(np.float32)
    im = (pt[rng.permutation(n)[:m]] + 0.4 * rng.standard_normal((m, d))).astype(np.float32)
    res = {}
    for tc in ("1", "0"):
        os.environ["NMR_GEMM_TC"] = tc
        MM._engines.clear()
        try:
            out = MM.coarse_match(torch.from_numpy(im).cuda(), torch.from_numpy(pt).cuda(), temperature=10.0, mutual=True, match_thres=0.2)
        finally:
            os.environ.pop("NMR_GEMM_TC")
        res[tc] = (out["conf_matrix"][0].cpu().numpy(), out["match_ids"][1].cpu().numpy(), out["match_ids"][2].cpu().numpy())
    MM._engines.clear()
    np.testing.assert_allclose(res["1"][0], res["0"][0], rtol=3e-4, atol=1e-9)
    assert np.array_equal(res["1"][1], res["0"][1]) and np.array_equal(res["1"][2], res["0"][2]) and len(res["1"][1]) > 0.8 * m
