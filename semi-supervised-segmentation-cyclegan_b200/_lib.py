"""ctypes binding of libsscg_b200.so (the C ABI declared in include/sscg_b200.h).

The library is built in-tree by csrc/build.sh (see __graft_entry__.build).  There is no fallback:
if the shared object is missing or a call fails, the caller gets an exception carrying
sscg_last_error().
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSCG_LIB", os.path.join(_HERE, "libsscg_b200.so"))   # SSCG_LIB: tuning builds only

SSCG_MAX_TAPS = 64
SSCG_LOSS_WS_BYTES = 16384
SSCG_STAT_WORDS = 2          # int64 words per plane-sum value (binned fixed-point accumulators, see the header)
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
PAD_NONE, PAD_ZERO, PAD_REFLECT = 0, 1, 2


class SscgError(RuntimeError):
    pass


class Tap(C.Structure):
    _fields_ = [("dh", C.c_int8), ("dw", C.c_int8), ("brow", C.c_int16)]


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("sN", C.c_int64), ("sH", C.c_int64), ("sW", C.c_int64)]


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", View), ("x_lo", C.c_void_p),
        ("stride", C.c_int32), ("Kc", C.c_int32), ("org_h", C.c_int32), ("org_w", C.c_int32),
        ("n_phases", C.c_int32), ("phase_start", C.c_int32 * 5), ("taps", Tap * SSCG_MAX_TAPS),
        ("w", C.c_void_p), ("w_lo", C.c_void_p), ("w_rows", C.c_int32), ("Co_pad", C.c_int32),
        ("split", C.c_int32),
        ("y", C.c_void_p), ("y_fp32", C.c_int32),
        ("y_sN", C.c_int64), ("y_sH", C.c_int64), ("y_sW", C.c_int64),
        ("y_oh", C.c_int32), ("y_ow", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32), ("slope", C.c_float),
        ("stats", C.c_void_p),
        ("TH", C.c_int32), ("TW", C.c_int32), ("BN", C.c_int32), ("tag", C.c_int32),
        ("shift_kw", C.c_int32), ("shift_brow_step", C.c_int32), ("shift_base_mode", C.c_int32),
        ("flat_pitch", C.c_int32), ("flat_hw", C.c_int32), ("flat_n", C.c_int32),
        ("rw_pitch", C.c_int32),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("dy", View), ("dy_lo", C.c_void_p), ("x", View), ("x_lo", C.c_void_p),
        ("stride", C.c_int32), ("Kc", C.c_int32), ("org_h", C.c_int32), ("org_w", C.c_int32),
        ("n_taps", C.c_int32), ("taps", Tap * SSCG_MAX_TAPS),
        ("Co_pad", C.c_int32), ("split", C.c_int32),
        ("dw", C.c_void_p), ("w_rows", C.c_int32),
        ("TH", C.c_int32), ("TW", C.c_int32), ("BN", C.c_int32), ("ksplit", C.c_int32), ("tag", C.c_int32),
        ("ws", C.c_void_p), ("rw_pitch", C.c_int32),
    ]


class Wgrad7Args(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dy", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("Cy", C.c_int32), ("dw", C.c_void_p), ("ws", C.c_void_p), ("tag", C.c_int32)]


class ApplyArgs(C.Structure):
    _fields_ = [
        ("raw", C.c_void_p), ("raw_fp32", C.c_int32),
        ("stats", C.c_void_p), ("eps", C.c_float),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("act", C.c_int32), ("slope", C.c_float),
        ("drop_seed", C.c_uint64), ("drop_ctr", C.c_void_p),
        ("res", View), ("res_lo", C.c_void_p),
        ("dst", C.c_void_p), ("dst_lo", C.c_void_p),
        ("pad", C.c_int32), ("pad_mode", C.c_int32),
    ]


class BwdArgs(C.Structure):
    _fields_ = [
        ("raw", C.c_void_p), ("raw_fp32", C.c_int32),
        ("stats", C.c_void_p), ("eps", C.c_float),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("act", C.c_int32), ("slope", C.c_float),
        ("drop_seed", C.c_uint64), ("drop_ctr", C.c_void_p),
        ("dyp", View), ("dyp_fp32", C.c_int32),
        ("pad", C.c_int32), ("pad_mode", C.c_int32),
        ("skip", View), ("skip_fp32", C.c_int32),
        ("g_out", C.c_void_p), ("g_fp32", C.c_int32),
        ("dz", C.c_void_p), ("dz_fp32", C.c_int32),
        ("dz_lo", C.c_void_p),
        ("bstats", C.c_void_p),
        ("dz_pad", C.c_int32),
        ("draw_pad", C.c_int32),
    ]


class WprepArgs(C.Structure):
    _fields_ = [
        ("w", C.c_void_p), ("transposed", C.c_int32),
        ("Co", C.c_int32), ("Ci", C.c_int32), ("KH", C.c_int32), ("KW", C.c_int32),
        ("mode", C.c_int32), ("Cp", C.c_int32), ("rows_pad", C.c_int32), ("Kc", C.c_int32),
        ("dst", C.c_void_p), ("dst_lo", C.c_void_p),
    ]


class Conv7Args(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_pitch", C.c_int32), ("N", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32),
                ("w", C.c_void_p), ("CoW", C.c_int32), ("n_ntiles", C.c_int32), ("ksteps", C.c_int32),
                ("c_store", C.c_int32), ("y", C.c_void_p), ("y_fp32", C.c_int32), ("y_sN", C.c_int64),
                ("y_sH", C.c_int64), ("y_sW", C.c_int64), ("bias", C.c_void_p), ("act", C.c_int32), ("tag", C.c_int32)]


class WbatchEntry(C.Structure):
    _fields_ = [("a", WprepArgs), ("slab", C.c_void_p), ("grad", C.c_void_p), ("start", C.c_int64)]


_SIGNATURES = {
    "sscg_conv_igemm": [C.POINTER(ConvArgs), C.c_void_p],
    "sscg_conv_wgrad": [C.POINTER(WgradArgs), C.c_void_p],
    "sscg_conv_wgrad_ws_bytes": [C.POINTER(WgradArgs)],
    "sscg_conv_wgrad_ctas_per_sm": [C.c_int32, C.c_int32],
    "sscg_set_pdl": [C.c_int32],
    "sscg_conv_wgrad7": [C.POINTER(Wgrad7Args), C.c_void_p],
    "sscg_conv_wgrad7_ws_bytes": [C.POINTER(Wgrad7Args)],
    "sscg_pack_nchw": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p],
    "sscg_unpack_fold": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                         C.c_int32, C.c_void_p, C.c_void_p],
    "sscg_bias_grad": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_float, C.c_void_p],
    "sscg_onehot_pack": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                         C.c_int32, C.c_int32, C.c_void_p],
    "sscg_unpack_nhwc": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    "sscg_in_apply": [C.POINTER(ApplyArgs), C.c_void_p],
    "sscg_in_bwd_prep": [C.POINTER(BwdArgs), C.c_void_p],
    "sscg_in_bwd_apply": [C.POINTER(BwdArgs), C.c_void_p, C.c_void_p, C.c_void_p],
    "sscg_set_stream_norm": [C.c_int32],
    "sscg_wprep": [C.POINTER(WprepArgs), C.c_void_p],
    "sscg_wgrad_unpack": [C.POINTER(WprepArgs), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p],
    "sscg_conv7_nexp": [C.POINTER(Conv7Args), C.c_void_p],
    "sscg_wprep_batch": [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p],
    "sscg_wgrad_unpack_batch": [C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p],
    "sscg_fill_zero": [C.c_void_p, C.c_int64, C.c_void_p],
    "sscg_seg_head_fwd": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p, C.c_void_p],
    "sscg_seg_head_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                          C.c_void_p, C.c_void_p],
    "sscg_lsgan_fwd": [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p],
    "sscg_lsgan_bwd": [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p],
    "sscg_l1_fwd": [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p],
    "sscg_l1_bwd": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    "sscg_adam_flat": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                       C.c_float, C.c_void_p, C.c_void_p],
    "sscg_interp_bilinear_fwd": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_void_p],
    "sscg_interp_bilinear_bwd": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p],
    "sscg_confusion": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p],
    "sscg_device_error": [],
    "sscg_version": [],
    "sscg_prof_begin": [C.c_uint32],
    "sscg_prof_end": [C.POINTER(C.c_float), C.POINTER(C.c_int32)],
}

_lib = None


def exported_symbols():
    """Names every entry point include/sscg_b200.h declares (used by the CPU-side ABI test)."""
    return list(_SIGNATURES.keys()) + ["sscg_last_error", "sscg_launch_count"]


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SscgError(
            f"{LIB_PATH} is missing: the sm_100a extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or csrc/build.sh). "
            "There is no CPU or PyTorch fallback for the CUDA path.")
    l = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int64 if name.endswith("_ws_bytes") else C.c_int
    l.sscg_last_error.restype = C.c_char_p
    l.sscg_last_error.argtypes = []
    l.sscg_launch_count.restype = C.c_uint64
    l.sscg_launch_count.argtypes = []
    _lib = l
    return l


def check(rc, what=""):
    if rc != 0:
        msg = lib().sscg_last_error().decode("utf-8", "replace")
        raise SscgError(f"{what} failed (rc={rc}): {msg}")


def make_view(ptr, N, H, W, Cc, sN, sH, sW):
    v = View()
    v.ptr = ptr
    v.N, v.H, v.W, v.C = N, H, W, Cc
    v.sN, v.sH, v.sW = sN, sH, sW
    return v
