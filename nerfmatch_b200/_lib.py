"""ctypes binding of libnmr.so (include/nmr.h).  No torch types cross this boundary: tensors are
passed as raw device pointers + sizes, the CUDA stream as an integer handle.

The library is built in-tree (``nerfmatch_b200/libnmr.so``) by ``build()`` / ``make -C
nerfmatch_b200/csrc``.  There is no fallback: if the shared object is missing or a call fails,
``NmrError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnmr.so")
ABI_VERSION = 2

NMR_EMBED_NORMAL, NMR_EMBED_MIP = 0, 1
NMR_NET_COARSE, NMR_NET_FINE = 0, 1
NMR_PREC_FP32, NMR_PREC_BF16, NMR_PREC_FP16 = 0, 1, 2
PRECISIONS = {"fp32": NMR_PREC_FP32, "bf16": NMR_PREC_BF16, "fp16": NMR_PREC_FP16}
NMR_RNG_NONE, NMR_RNG_PHILOX = 0, 1
NMR_FLAG_RET_PFEAT, NMR_FLAG_FEAT_MAX, NMR_FLAG_WHITE_BG, NMR_FLAG_SKIP_COARSE_FEAT, NMR_FLAG_PERTURB = 1, 2, 4, 8, 16

_fp = C.POINTER(C.c_float)


class NmrError(RuntimeError):
    pass


class NmrConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("embed_type", C.c_int32), ("xyz_num_freqs", C.c_int32),
        ("dirs_num_freqs", C.c_int32), ("num_pts_coarse", C.c_int32), ("num_pts_fine", C.c_int32),
        ("hid_dim", C.c_int32), ("layer_num", C.c_int32), ("skip_layer", C.c_int32), ("app_dim", C.c_int32),
        ("stop_layer_coarse", C.c_int32), ("stop_layer_fine", C.c_int32), ("use_disp", C.c_int32),
        ("precision", C.c_int32), ("mip_var_scale", C.c_float), ("resample_padding", C.c_float),
        ("max_chunk_points", C.c_int64),
    ]


class NmrWeightPtrs(C.Structure):
    _fields_ = [
        ("pts_w", C.c_void_p * 8), ("pts_b", C.c_void_p * 8),
        ("views_w", C.c_void_p), ("views_b", C.c_void_p), ("feature_w", C.c_void_p), ("feature_b", C.c_void_p),
        ("alpha_w", C.c_void_p), ("alpha_b", C.c_void_p), ("rgb_w", C.c_void_p), ("rgb_b", C.c_void_p),
    ]


class NmrRenderArgs(C.Structure):
    _fields_ = [
        ("rays", C.c_void_p), ("n", C.c_int64), ("ncols", C.c_int32), ("flags", C.c_uint32),
        ("ray_id", C.c_void_p), ("t_rand", C.c_void_p), ("u_jitter", C.c_void_p),
        ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("unnorm", C.c_void_p),
        ("feat_coarse", C.c_void_p), ("feat_fine", C.c_void_p), ("pts_coarse", C.c_void_p), ("pts_fine", C.c_void_p),
        ("rgb_coarse", C.c_void_p), ("rgb_fine", C.c_void_p), ("depth_coarse", C.c_void_p), ("depth_fine", C.c_void_p),
        ("pt3d", C.c_void_p),
        ("z_coarse", C.c_void_p), ("z_fine", C.c_void_p), ("weights_coarse", C.c_void_p), ("weights_fine", C.c_void_p),
        ("raw_coarse", C.c_void_p), ("raw_fine", C.c_void_p), ("idx_fine", C.c_void_p),
        ("sigma_noise_coarse", C.c_void_p), ("sigma_noise_fine", C.c_void_p), ("s_fine", C.c_void_p),
    ]


# every symbol include/nmr.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
SYMBOLS = {
    "nmr_abi_version": (C.c_int, []),
    "nmr_create": (C.c_int, [C.POINTER(NmrConfig), C.POINTER(_vp)]),
    "nmr_destroy": (None, [_vp]),
    "nmr_last_error": (C.c_char_p, [_vp]),
    "nmr_set_weights": (C.c_int, [_vp, C.c_int, C.POINTER(NmrWeightPtrs)]),
    "nmr_set_appearance": (C.c_int, [_vp, _vp, _i32]),
    "nmr_set_precision": (C.c_int, [_vp, C.c_int]),
    "nmr_raygen": (C.c_int, [_vp, _fp, _fp, _i32, _i32, _i32, _vp, C.POINTER(_i32), _vp]),
    "nmr_render": (C.c_int, [_vp, C.POINTER(NmrRenderArgs), _vp]),
    "nmr_render_host": (C.c_int, [_vp, C.POINTER(NmrRenderArgs), _vp]),
    "nmr_render_view": (C.c_int, [_vp, _fp, _fp, _i32, _i32, _i32, C.POINTER(NmrRenderArgs), _vp]),
    "nmr_stage_sample_coarse": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "nmr_stage_resample": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nmr_stage_encode": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _vp, _vp]),
    "nmr_stage_mlp": (C.c_int, [_vp, C.c_int, _vp, _i64, _i32, _vp, _vp, _vp]),
    "nmr_stage_composite": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _u32] + [_vp] * 10),
    "nmr_stage_encode_backward": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _vp, _vp, _vp]),
    "nmr_stage_mlp_backward": (C.c_int, [_vp, C.c_int, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nmr_stage_mlp_masks": (C.c_int, [_vp, C.c_int, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "nmr_stage_composite_backward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nmr_peer_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp), _vp]),
    "nmr_peer_free": (C.c_int, [_vp, _vp]),
    "nmr_peer_open": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "nmr_peer_close": (C.c_int, [_vp, _vp]),
    "nmr_peer_signal": (C.c_int, [_vp, _vp, _u32, _vp]),
    "nmr_peer_wait": (C.c_int, [_vp, _vp, _i32, _u32, _vp]),
    "nmr_peer_wait_status": (C.c_int, [_vp]),
    "nmr_coarse_match": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, C.c_float, _i32, _i32, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nmr_stage_mlp_probe": (C.c_int, [_vp, C.c_int, _vp, _i64, _i32, _vp, _vp]),
    "nmr_debug_trap_code": (C.c_int, []),
    "nmr_set_tc_cluster": (C.c_int, [C.c_int]),
    "nmr_debug_timeline": (C.c_int, [C.POINTER(C.c_longlong), C.c_int]),
    "nmr_launch_count": (_i64, [_vp]),
    "nmr_ray_dirs": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "nmr_rays_c2w": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "nmr_prepare_rays": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_float, _vp, C.c_float, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "nmr_copy_host_staging": (C.c_int, [_vp, C.c_int32, _vp, C.c_size_t, _vp]),
    "nmr_set_timing": (C.c_int, [_vp, C.c_int]),
    "nmr_set_fuse_composite": (C.c_int, [_vp, C.c_int]),
    "nmr_last_stage_ms": (C.c_int, [_vp, _fp, _fp, _fp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libnmr.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise NmrError("building libnmr.so failed (see output above)")
    return LIB_PATH


def load():
    """dlopen libnmr.so and type every entry point.  Raises NmrError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NmrError(
            f"{LIB_PATH} not found: build it with `make -C nerfmatch_b200/csrc` or __graft_entry__.build(); "
            "there is no CPU / PyTorch fallback for the render path"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.nmr_abi_version() != ABI_VERSION:
        raise NmrError("libnmr.so ABI version mismatch: rebuild it")
    _lib = lib
    return lib


def check(handle, status: int, what: str):
    if status != 0:
        msg = load().nmr_last_error(handle)
        raise NmrError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")
