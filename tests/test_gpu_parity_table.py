"""Parity at the BASELINE config sizes and on every reference golden, in all three precisions, with the bounds DESIGN.md
section 2 states (measured table: profiles/r02_parity.json, produced by tests/parity_table.py).

  fp32 path      the north star everywhere: feature / depth rel-err < 1e-3, cosine ~ 1, 3D points < 1 mm (scene scale 3 m)
  bf16 / fp16    (a) north-star cosine >= 0.999 on every ray of the config-size views (19 200 rays each) and of the goldens
                     whose field 16-bit operands can resolve; un-normalised points < 1 mm in mip mode
                 (b) on every golden: within 3x of the error the EMULATED 16-bit-operand oracle makes against the same
                     reference (oracle.mlp_forward_rounded: the reference algorithm with operands rounded, fp32 accumulate)
                     - i.e. the CUDA path is as close to the reference as 16-bit tensor-core operands allow; the residual
                     is the network's own sensitivity (sigma gains of ~5e3, 2^14-frequency aliasing of an untrained classic
                     NeRF, arg-max features), not the kernel
                 (c) absolute per-case bounds from the measured table, with head-room, so a regression shows up
"""
import json
import os

import pytest

torch = pytest.importorskip("torch")

import parity_table as PT  # noqa: E402

pytestmark = pytest.mark.gpu

STRICT_16 = ("mip64", "mip128_peaky", "normal_64_128")          # goldens on which 16-bit operands keep cosine >= 0.999 on every ray
# measured (profiles/r02_parity.json) -> asserted: fraction of rays with cosine >= 0.999, median |d pt3d| in metres
BOUNDS = {
    "bf16": {"mip64": (1.0, 2e-3), "mip128_peaky": (1.0, 5e-3), "normal_64_128": (1.0, 4e-2), "normal_64_128_peaky": (0.60, 1.5e-2),
             "mip64_app_max_white": (0.70, 2e-3), "normal_app_stop5": (0.55, 6e-3)},
    "fp16": {"mip64": (1.0, 2.5e-4), "mip128_peaky": (1.0, 7e-4), "normal_64_128": (1.0, 1.5e-2), "normal_64_128_peaky": (0.93, 4e-3),
             "mip64_app_max_white": (0.95, 2.5e-4), "normal_app_stop5": (0.75, 1.5e-3)},
}


@pytest.fixture(scope="module")
def table():
    rows = PT.golden_rows()
    rows.update(PT.view_rows(specs={"config1_120x160/normal": dict(embed="normal", ds=4), "config1_120x160/mip": dict(embed="mip", ds=4)}))
    try:
        out = os.path.join(PT.ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        json.dump(rows, open(os.path.join(out, "parity_table_from_test.json"), "w"), indent=1)
    except OSError:
        pass
    return rows


def test_fp32_path_meets_the_north_star_everywhere(table):
    for key, r in table.items():
        if not key.endswith("/fp32"):
            continue
        lin = r.get("feat_comb", "lin") == "lin"
        for k in ("coarse", "fine"):
            assert r[f"feat_{k}"]["cos_min"] >= 0.99999, (key, k)
            if lin:
                assert r[f"feat_{k}"]["rel_l2"] < 1e-3, (key, k)
            assert r[f"depth_{k}"]["rel_l2"] < 1e-3, (key, k)
        assert r["pt3d"]["metres_max"] < 1e-3, (key, r["pt3d"])                  # 3D points within 1 mm
        if "idx_fine_equal" in r:
            assert r["idx_fine_equal"] > 0.9995, key                             # (bit-exact given equal weights: test_gpu_parity.py)


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_config_size_views_16bit(table, prec):
    for embed in ("normal", "mip"):
        r = table[f"config1_120x160/{embed}/{prec}"]
        assert r["n_rays"] == 19200
        for k in ("coarse", "fine"):
            assert r[f"feat_{k}"]["cos_min"] >= (0.999 if prec == "bf16" else 0.9999), (embed, k, r[f"feat_{k}"])   # north star, every ray
        if embed == "mip":       # the shipped configurations: 1 mm holds in both formats
            assert r["pt3d"]["metres_max"] < (1e-3 if prec == "bf16" else 1e-4), r["pt3d"]
            assert r["depth_fine"]["rel_max"] < (3e-3 if prec == "bf16" else 5e-4)
        else:                    # classic NeRF, untrained: every coarse-weight change is amplified by the 2^14-frequency field
            assert r["pt3d"]["metres_p99"] < (6e-3 if prec == "bf16" else 1e-3), r["pt3d"]
            assert r["pt3d"]["metres_max"] < (1.5e-2 if prec == "bf16" else 5e-3), r["pt3d"]
            assert r["depth_fine"]["rel_l2"] < (8e-3 if prec == "bf16" else 1.2e-3)
        assert r["sigma_coarse"]["rel_l2"] < (3e-3 if prec == "bf16" else 4e-4)   # 16-bit operand rounding through 8 layers


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_goldens_16bit_against_reference_and_operand_envelope(table, prec):
    for name in PT.RENDER_CASES:
        r = table[f"golden/{name}/{prec}"]
        emu = table[f"golden/{name}/{prec}_emulated_oracle_vs_reference"]
        frac, med_pt = BOUNDS[prec][name]
        ff = r["feat_fine"]
        if name in STRICT_16:
            assert ff["cos_min"] >= 0.999 and r["feat_coarse"]["cos_min"] >= 0.999, (name, ff)
        assert ff["frac_cos_ge_0999"] >= frac, (name, ff)
        assert ff["cos_median"] >= (0.9994 if prec == "bf16" else 0.99995), (name, ff)
        assert r["pt3d"]["metres_median"] <= med_pt, (name, r["pt3d"])
        # envelope: what 16-bit operands cost by themselves (medians: the small goldens' extremes are single rays)
        assert 1 - ff["cos_median"] <= 3 * (1 - emu["feat_fine"]["cos_median"]) + 2e-6, (name, ff, emu["feat_fine"])
        assert 1 - ff["cos_p1"] <= 3 * (1 - emu["feat_fine"]["cos_p1"]) + 2e-5, (name, ff, emu["feat_fine"])
        assert r["pt3d"]["metres_median"] <= 3 * emu["pt3d"]["metres_median"] + 2e-5, (name, r["pt3d"], emu["pt3d"])
        assert r["depth_fine"]["rel_median"] <= 3 * emu["depth_fine"]["rel_median"] + 1e-5, name
        assert r["sigma_coarse"]["rel_l2"] <= 1.25 * emu["sigma_coarse"]["rel_l2"], (name, r["sigma_coarse"], emu["sigma_coarse"])
        # fp16 operands carry 3 more significand bits than bf16: the density error drops accordingly
    for name in PT.RENDER_CASES:
        if prec == "fp16":
            assert table[f"golden/{name}/fp16"]["sigma_coarse"]["rel_l2"] < 0.2 * table[f"golden/{name}/bf16"]["sigma_coarse"]["rel_l2"], name
