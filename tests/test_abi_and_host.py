"""CPU-side checks: the C-ABI library loads and exports every symbol include/nmr.h declares, the
ctypes structs match the header, and the host logic of the shim (constructor, state_dict keys,
error behaviour without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from nmr_helpers import ref_style_config  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_text():
    return open(os.path.join(ROOT, "include", "nmr.h")).read()


def test_library_exports_every_declared_symbol():
    from nerfmatch_b200 import _lib as L

    declared = set(re.findall(r"NMR_API\s+[\w\s\*]+?\b(nmr_\w+)\s*\(", header_text()))
    assert len(declared) >= 19
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nmr_abi_version() == L.ABI_VERSION


def test_ctypes_structs_match_header_fields():
    from nerfmatch_b200 import _lib as L

    txt = header_text()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), txt, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return [re.sub(r"\[\d+\]", "", m) for m in re.findall(r"(\w+(?:\[\d+\])?)\s*;", body)]

    assert fields("nmr_config") == [f[0] for f in L.NmrConfig._fields_]
    assert fields("nmr_weight_ptrs") == [f[0] for f in L.NmrWeightPtrs._fields_]
    assert fields("nmr_render_args") == [f[0] for f in L.NmrRenderArgs._fields_]
    # sizes / a late field offset as the C compiler lays them out
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include <stddef.h>\n#include "nmr.h"\nint main(){printf("%zu %zu %zu %zu %zu", sizeof(nmr_config), ' \
          'sizeof(nmr_weight_ptrs), sizeof(nmr_render_args), offsetof(nmr_render_args, idx_fine), offsetof(nmr_config, max_chunk_points));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "s.c"), "-o", os.path.join(td, "s")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(td, "s")]).split()]
    assert sizes == [C.sizeof(L.NmrConfig), C.sizeof(L.NmrWeightPtrs), C.sizeof(L.NmrRenderArgs),
                     L.NmrRenderArgs.idx_fine.offset, L.NmrConfig.max_chunk_points.offset]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    from nerfmatch_b200 import NerfRenderer, NmrError
    from nerfmatch_b200 import _lib as L

    lib = L.load()
    cfg = L.NmrConfig(abi_version=L.ABI_VERSION, embed_type=1, xyz_num_freqs=15, dirs_num_freqs=4, num_pts_coarse=64, num_pts_fine=128,
                      hid_dim=256, layer_num=8, skip_layer=4, app_dim=0, stop_layer_coarse=-1, stop_layer_fine=3)
    h = C.c_void_p()
    assert lib.nmr_create(C.byref(cfg), C.byref(h)) == -2 and not h.value  # NMR_ERR_CUDA
    assert b"no CPU fallback" in lib.nmr_last_error(None)
    ren = NerfRenderer(ref_style_config("mip"), stop_layer=3, training=False)
    with pytest.raises(RuntimeError):
        ren.predict(torch.zeros(4, 12), 2, 2)


def test_create_rejects_bad_configs():
    from nerfmatch_b200 import _lib as L

    lib = L.load()
    h = C.c_void_p()
    bad_abi = L.NmrConfig(abi_version=99)
    assert lib.nmr_create(C.byref(bad_abi), C.byref(h)) == -1
    wide = L.NmrConfig(abi_version=L.ABI_VERSION, embed_type=1, xyz_num_freqs=15, dirs_num_freqs=4, num_pts_coarse=64, num_pts_fine=128,
                       hid_dim=512, layer_num=8, skip_layer=4)
    assert lib.nmr_create(C.byref(wide), C.byref(h)) == -4  # NMR_ERR_UNSUPPORTED
    many = L.NmrConfig(abi_version=L.ABI_VERSION, embed_type=0, xyz_num_freqs=15, dirs_num_freqs=4, num_pts_coarse=192, num_pts_fine=128,
                       hid_dim=256, layer_num=8, skip_layer=4)
    assert lib.nmr_create(C.byref(many), C.byref(h)) == -4
    assert lib.nmr_create(None, C.byref(h)) == -1


@pytest.mark.parametrize("embed,app,nkeys", [("mip", False, 50), ("normal", False, 48), ("mip", True, 51)])
def test_state_dict_keys_match_reference(embed, app, nkeys):
    """50 tensors (mip) / 48 (normal) under the reference's key names."""
    from nerfmatch_b200 import NerfRenderer
    from oracle.weights import make_state_dict

    ren = NerfRenderer(ref_style_config(embed, app=app), num_frames=7 if app else None, training=False, stop_layer=3)
    sd = make_state_dict(0, embed_type=embed, app_dim=16 if app else 0, num_frames=7 if app else 0)
    assert set(ren.state_dict().keys()) == set(sd.keys())
    assert len(sd) == nkeys
    for k, v in ren.state_dict().items():
        assert tuple(v.shape) == sd[k].shape, k
    ren.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert ren.nerf_coarse.stop_layer == -1 and ren.nerf_fine.stop_layer == 3  # renderer.py:99
    assert ren.ret_pfeat is False and ren.feat_comb == "lin" and ren.pfeat_mask is None
    nparam = sum(p.numel() for n, p in ren.named_parameters() if "scales" not in n and "embedding_a" not in n)
    assert nparam == (1219336 if embed == "mip" else 1222408) + (2 * 16 * 128 if app else 0)


def test_unsupported_reference_options_fail_loudly():
    from nerfmatch_b200 import NerfRenderer

    cfg = ref_style_config("mip")
    cfg.data.out_scr = True
    with pytest.raises(NotImplementedError):
        NerfRenderer(cfg)
    cfg = ref_style_config("mip")
    cfg.render.single_model = True                 # built: the fine net renders both passes; no nerf_coarse.* keys (renderer.py:76-87)
    ren = NerfRenderer(cfg, stop_layer=3)
    assert not hasattr(ren, "nerf_coarse") and not any(k.startswith("nerf_coarse") for k in ren.state_dict())
    cfg = ref_style_config("mip")
    cfg.coarse_nerf.skips = [3]
    with pytest.raises(NotImplementedError):
        NerfRenderer(cfg)


def test_scene_cache_format_both_directions(tmp_path):
    """the repo's reader loads a file written the way the reference writes it
    (tests/golden/ref_cache, produced by oracle/gen_golden.py::cache_cases with the reference's np.save line), and a file
    written by the repo's writer was loaded by the REFERENCE's own load_frame_3d (what it returned is in the fixture)."""
    from nerfmatch_b200.cache import frame_cache_name, load_frame_3d, write_frame
    from nmr_helpers import GOLDEN, load_golden

    g = load_golden("cache_format.npz")
    frame = {"file_path": "seq-01/frame-000007.color.png"}
    assert frame_cache_name(frame["file_path"]) == "seq-01_frame-000007"
    # reference-written file -> our reader
    pt3d, pt_feat, mask, un = load_frame_3d(frame, os.path.join(GOLDEN, "ref_cache"))
    np.testing.assert_array_equal(pt3d, g["pt3d"])
    np.testing.assert_array_equal(pt_feat, g["pt_feat"])
    np.testing.assert_array_equal(un, g["unnorm_scene"])
    assert mask.dtype == np.bool_ and mask.all() and len(mask) == len(pt3d)
    ref_item = np.load(os.path.join(GOLDEN, "ref_cache", "seq-01_frame-000007.npy"), allow_pickle=True).item()
    # our writer -> same keys / dtypes / values as the reference-written dict, and what the reference's reader returned
    path = write_frame(str(tmp_path), "seq-01_frame-000007", g["pt3d"], g["unnorm_scene"], g["pt_feat"], g["pt_color_raw"])
    item = np.load(path, allow_pickle=True).item()
    assert sorted(item) == sorted(ref_item) == list(g["ref_written.keys"]) == list(g["ref_reads_ours.keys"])
    for k in item:
        assert item[k].dtype == ref_item[k].dtype and item[k].shape == ref_item[k].shape, k
        np.testing.assert_array_equal(item[k], ref_item[k])
    np.testing.assert_array_equal(g["ref_reads_ours.pt3d"], g["pt3d"])
    np.testing.assert_array_equal(g["ref_reads_ours.pt_feat"], g["pt_feat"])
    np.testing.assert_array_equal(g["ref_reads_ours.unnorm_scene"], g["unnorm_scene"])
    assert g["ref_reads_ours.mask"].all() and bool(g["ref_reads_ours.c2w_is_none"])
    a = load_frame_3d(frame, str(tmp_path), return_pose=True)
    assert len(a) == 5 and a[4] is None
    np.testing.assert_array_equal(a[1], g["pt_feat"])
