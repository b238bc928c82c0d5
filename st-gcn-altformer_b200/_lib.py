"""ctypes binding of lib/libaltformer_b200.so (the C ABI declared in include/altformer_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not an sm_100
part, importing/using the ops raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libaltformer_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_RELU = 0, 1, 2, 3
RS_VALUE, RS_BIAS = 0, 1
GCN0_NMOM, GCN0_NSTAT_BASE, GCN0_SLOTS = 96, 160, 32

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmTn(C.Structure):
    _fields_ = [("A", vp), ("B", vp), ("C", vp), ("C2", vp), ("rows_per_batch", i64), ("batches", i32), ("N", i32),
                ("k_per_tap", i32), ("taps", i32), ("tap_row_stride", i32), ("tap_pad", i32), ("lda", i32),
                ("ldb", i32), ("ldc", i32), ("b_mn_major", i32), ("out_dtype", i32), ("act", i32), ("alpha", f32),
                ("bias", vp), ("pos", vp), ("pos_rows", i32), ("aux", vp), ("aux_dtype", i32), ("ldaux", i32),
                ("residual", vp), ("res_dtype", i32), ("ldres", i32), ("row_scale", vp), ("row_scale_div", i32),
                ("row_scale_mode", i32)]


class GemmDw(C.Structure):
    _fields_ = [("G", vp), ("X", vp), ("dW", vp), ("rows_per_batch", i64), ("batches", i32), ("N1", i32), ("N2", i32),
                ("ldg", i32), ("ldx", i32), ("ld1", i64), ("ld2", i64), ("x_row_shift", i32), ("alpha", f32), ("dbias", vp),
                ("taps", i32), ("tap_row_stride", i32), ("tap_dw_stride", i64),
                ("dbias_row_scale", vp), ("row_scale_div", i32)]


class GemmSimt(C.Structure):
    _fields_ = [("A", vp), ("B", vp), ("C", vp), ("bias", vp), ("M", i32), ("N", i32), ("K", i32), ("sai", i64),
                ("sak", i64), ("sbj", i64), ("sbk", i64), ("sci", i64), ("scj", i64), ("a_dtype", i32),
                ("b_dtype", i32), ("c_dtype", i32), ("alpha", f32), ("beta", f32)]


class Gcn0Fwd(C.Structure):
    _fields_ = [("x", vp), ("A", vp), ("PA", vp), ("Wa", vp * 3), ("ba", vp * 3), ("Wb", vp * 3), ("bb", vp * 3),
                ("Wd", vp * 3), ("bd", vp * 3), ("Wdn", vp), ("bdn", vp), ("bn_g", vp), ("bn_b", vp), ("dn_g", vp),
                ("dn_b", vp), ("bn_rm", vp), ("bn_rv", vp), ("dn_rm", vp), ("dn_rv", vp), ("N", i32), ("T", i32),
                ("V", i32), ("Cout", i32), ("IC", i32), ("training", i32), ("momentum", f32), ("eps", f32),
                ("Mmat", vp), ("moments", vp), ("counter", vp), ("stats", vp), ("Wfold", vp), ("Aop", vp), ("colsum", vp), ("Wfrag", vp), ("y", vp), ("y_dtype", i32),
                ("precise", i32)]


class Gcn0Bwd(C.Structure):
    _fields_ = [("f", Gcn0Fwd), ("dy", vp), ("ws", vp), ("dPA", vp), ("dWa", vp * 3), ("dba", vp * 3),
                ("dWb", vp * 3), ("dbb", vp * 3), ("dWd", vp * 3), ("dbd", vp * 3), ("dWdn", vp), ("dbdn", vp),
                ("dbn_g", vp), ("dbn_b", vp), ("ddn_g", vp), ("ddn_b", vp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python st-gcn-altformer_b200/build.py` "
                "(altformer_b200 has no CPU or eager fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.afb_last_error.restype = C.c_char_p
        _declare(_lib)
    return _lib


def _declare(L):
    P = C.POINTER
    sig = {
        "afb_version": [],
        "afb_device_ok": [i32],
        "afb_gemm_tn": [P(GemmTn), vp],
        "afb_gemm_dw": [P(GemmDw), vp],
        "afb_gemm_simt": [P(GemmSimt), vp],
        "afb_cast": [vp, i32, vp, i32, i64, vp],
        "afb_copy2d": [vp, i32, i64, vp, i32, i64, i64, i32, vp],
        "afb_cast_transpose": [vp, vp, i32, i32, vp],
        "afb_conv_weight_pack": [vp, vp, vp, i32, i32, i32, vp],
        "afb_conv_dw_unpack": [vp, vp, i32, i32, i32, vp],
        "afb_split3": [vp, vp, i64, i32, i32, vp],
        "afb_layernorm_fwd": [vp, i32, vp, vp, vp, i32, vp, vp, i64, i32, f32, vp],
        "afb_layernorm_bwd": [vp, i32, vp, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, i64, i32, vp],
        "afb_attention_fwd": [vp, vp, i32, i64, i32, i32, i32, f32, vp, vp],
        "afb_attention_bwd": [vp, vp, vp, i32, i64, i32, i32, i32, f32, vp],
        "afb_colstats": [vp, i32, i64, i32, i32, vp, vp, vp],
        "afb_colsum": [vp, i32, i64, i32, i32, vp, i32, vp, vp],
        "afb_bn_finalize": [vp, vp, i64, i32, vp, vp, vp, vp, f32, f32, i32, vp, vp, vp, vp, vp],
        "afb_bn_act_fwd": [vp, i32, vp, vp, vp, vp, i32, i32, vp, vp, i32, i64, i32, i32, i32, vp],
        "afb_bn_bwd_reduce": [vp, vp, i32, vp, i32, vp, i32, vp, vp, vp, vp, i32, vp, vp, i64, i32, i32, i32, vp],
        "afb_bn_bwd_apply": [vp, vp, i32, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, i32, i64, i32,
                             i32, i32, vp],
        "afb_pool_mean_fwd": [vp, vp, i32, i64, i32, i32, vp],
        "afb_pool_mean_bwd": [vp, vp, i32, i64, i32, i32, vp],
        "afb_pool_max_fwd": [vp, vp, vp, i32, i64, i32, i32, vp],
        "afb_pool_max_bwd": [vp, vp, vp, i32, i64, i32, i32, vp],
        "afb_softmax_ce": [vp, vp, vp, vp, i32, i32, vp],
        "afb_adamw": [vp, vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, f32, vp],
        "afb_step_inc": [vp, vp],
        "afb_scale_rows": [vp, vp, i32, i64, i32, vp, i32, vp],
        "afb_gcn0_fwd": [P(Gcn0Fwd), vp],
        "afb_gcn0_aop_bytes": [i32, i32],
        "afb_gcn0_fused_stamps": [vp, i32],
        "afb_gcn0_bwd": [P(Gcn0Bwd), vp],
        "afb_agcn_scores_fwd": [vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "afb_agcn_aggregate_fwd": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "afb_agcn_aggregate_bwd": [vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, vp],
        "afb_agcn_scores_bwd": [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "afb_agcn_aggregate_fwd_mma": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "afb_agcn_aggregate_bwd_mma": [vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, vp],
        "afb_agcn_scores_fwd_mma": [vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "afb_agcn_scores_bwd_mma": [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "afb_bone_stream": [vp, vp, vp, i64, i32, vp],
        "afb_motion_stream": [vp, vp, i32, i32, i32, vp],
        "afb_palm_center": [vp, vp, i32, i32, i32, i32, vp],
        "afb_augment": [vp, vp, i64, i32, i32, vp, vp, vp],
        "afb_mul": [vp, vp, vp, i32, i64, vp],
        "afb_frame_phase": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
        "afb_axpby": [vp, f32, vp, f32, vp, i64, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int64 if name == "afb_gcn0_aop_bytes" else C.c_int
    L._afb_signatures = sig


EXPORTS = None


def exported_names():
    """Every symbol include/altformer_b200.h declares (parsed from the header itself)."""
    import re
    hdr = os.path.join(HERE, "..", "include", "altformer_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(afb_[a-z0-9_]+)\s*\(", txt)))


def check(rc, what=""):
    if rc != 0:
        msg = lib().afb_last_error().decode(errors="replace")
        raise RuntimeError(f"altformer_b200 {what} failed (code {rc}): {msg}")
