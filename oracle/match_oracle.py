"""numpy oracle of NeRFMatch-Mini coarse matching  --  TEST INFRASTRUCTURE ONLY (see oracle/nerf_oracle.py header).

Restates nerfmatch/nerfmatch_coarse_trainer.py:83 (norm_feat), :223-234 (coarse_matching) and
nerfmatch/modules/extract_matches.py:13-40 (extract_mutual_matches, inference branch) for batch size 1.
Pinned against fixtures produced by the reference functions themselves (oracle/gen_golden.py -> tests/golden/match.npz)."""
import numpy as np

F32 = np.float32


def norm_feat(x):
    return (x / (np.sqrt(np.sum(x * x, axis=-1, keepdims=True, dtype=F32)) + F32(1e-6))).astype(F32)


def softmax(x, axis):
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def coarse_matching(im_feat, pt_feat, temperature=10.0, temp_type="mul", im_mask=None, pt_mask=None):
    """im_feat [m,d], pt_feat [n,d] -> conf [m,n] = softmax over m * softmax over n of the tempered cosine similarity."""
    im, pt = norm_feat(im_feat.astype(F32)), norm_feat(pt_feat.astype(F32))
    sim = (im @ pt.T).astype(F32)
    sim = sim / F32(temperature) if temp_type == "div" else sim * F32(temperature)
    imk = np.ones(len(im), bool) if im_mask is None else im_mask.astype(bool)
    ptk = np.ones(len(pt), bool) if pt_mask is None else pt_mask.astype(bool)
    sim = np.where(imk[:, None] & ptk[None, :], sim, F32(-1e9)).astype(F32)
    return (softmax(sim, 0) * softmax(sim, 1)).astype(F32)


def extract_mutual_matches(conf, mutual=True, threshold=0.0):
    mask = conf > threshold
    max_w = conf == conf.max(axis=1, keepdims=True)
    max_h = conf == conf.max(axis=0, keepdims=True)
    mask = mask & max_w & max_h if mutual else mask & max_w
    mask_v = mask.max(axis=1)
    all_j = mask.argmax(axis=1)          # first True, like torch's max over a bool tensor
    i_ids = np.where(mask_v)[0]
    j_ids = all_j[i_ids]
    return i_ids.astype(np.int64), j_ids.astype(np.int64), conf[i_ids, j_ids]
