"""Functional host layer: torch tensors in, C-ABI kernel launches on the current CUDA stream out.

torch is used for device memory (torch.empty), the stream handle and nothing else; every byte of
arithmetic happens in lib/libaltformer_b200.so.  All functions require CUDA tensors and raise
otherwise (no CPU / eager fallback).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_GELU_BWD, ACT_NONE, ACT_RELU, BF16, F32, RS_BIAS, RS_VALUE  # noqa: F401

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"altformer_b200: unsupported dtype {t.dtype} (float32 or bfloat16)")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("altformer_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError(f"altformer_b200: non-contiguous tensor of shape {tuple(t.shape)}")


_checked_devices = set()


def ensure_device(t):
    d = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if d not in _checked_devices:
        _lib.check(_lib.lib().afb_device_ok(d), "device check")
        _checked_devices.add(d)


LAUNCHES = [0]  # kernels launched through the C ABI (bench.py reports launches per step)
_KERNELS_PER_CALL = {"afb_gcn0_fwd": 2, "afb_gcn0_bwd": 4}


def _call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args), name)
    LAUNCHES[0] += _KERNELS_PER_CALL.get(name, 1)


# ------------------------------------------------------------------------------------------------
# GEMMs
# ------------------------------------------------------------------------------------------------
def gemm_tn(a, b, N, *, k_per_tap=None, taps=1, tap_row_stride=0, tap_pad=0, rows_per_batch=None, batches=1,
            b_mn_major=False, out_dtype=torch.bfloat16, bias=None, act=ACT_NONE, want_preact=False, aux=None,
            residual=None, row_scale=None, row_scale_div=1, row_scale_bias_only=False, pos=None, alpha=1.0, out=None):
    """C = epilogue(A @ B^T).  a [M, lda] bf16, b [N, K] bf16 (or [K, N] when b_mn_major).
    row_scale_bias_only: the rows of `a` already carry row_scale, so only the bias term is multiplied by it."""
    need_cuda(a, b, bias, aux, residual, row_scale, pos, out)
    ensure_device(a)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise RuntimeError("gemm_tn operands must be bfloat16")
    M, lda = a.shape
    if rows_per_batch is None:
        rows_per_batch, batches = M, 1
    assert rows_per_batch * batches == M
    if k_per_tap is None:
        k_per_tap = lda
    c = out if out is not None else torch.empty((M, N), device=a.device, dtype=out_dtype)
    c2 = torch.empty_like(c) if want_preact else None
    p = _lib.GemmTn(A=ptr(a), B=ptr(b), C=ptr(c), C2=ptr(c2), rows_per_batch=rows_per_batch, batches=batches, N=N,
                    k_per_tap=k_per_tap, taps=taps, tap_row_stride=tap_row_stride, tap_pad=tap_pad, lda=lda,
                    ldb=b.shape[1], ldc=N, b_mn_major=int(b_mn_major), out_dtype=dt(c), act=act, alpha=alpha,
                    bias=ptr(bias), pos=ptr(pos), pos_rows=0 if pos is None else pos.shape[-2],
                    aux=ptr(aux), aux_dtype=0 if aux is None else dt(aux), ldaux=0 if aux is None else aux.shape[1],
                    residual=ptr(residual), res_dtype=0 if residual is None else dt(residual),
                    ldres=0 if residual is None else residual.shape[1], row_scale=ptr(row_scale),
                    row_scale_div=row_scale_div, row_scale_mode=RS_BIAS if row_scale_bias_only else RS_VALUE)
    _call("afb_gemm_tn", C.byref(p), stream())
    return (c, c2) if want_preact else c


def gemm_dw(g, x, dW, *, N1=None, N2=None, rows_per_batch=None, batches=1, ld1=None, ld2=1, x_row_shift=0, alpha=1.0,
            g_col0=0, x_col0=0, dbias=None, dbias_row_scale=None, row_scale_div=1, taps=1, tap_row_stride=0, tap_dw_stride=0):
    """dW[n1*ld1 + n2*ld2] += alpha * sum_m g[m, g_col0 + n1] * x[m + shift, x_col0 + n2]  (fp32 atomics);
    dbias[n1] += alpha * sum_m (rs[m // row_scale_div]) * g[m, g_col0 + n1] when given (rs = dbias_row_scale or 1).
    taps in (2, 3): the launch produces `taps` gradients sharing g -- tap t reads x rows shifted by a further
    t * tap_row_stride and accumulates into dW.view(-1)[t * tap_dw_stride:] (k x 1 temporal conv)."""
    need_cuda(g, x, dW, dbias, dbias_row_scale)
    if g.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dW.dtype != torch.float32:
        raise RuntimeError("gemm_dw: g, x must be bfloat16 and dW float32")
    M = g.shape[0]
    if rows_per_batch is None:
        rows_per_batch, batches = M, 1
    N1 = g.shape[1] if N1 is None else N1
    N2 = x.shape[1] if N2 is None else N2
    p = _lib.GemmDw(G=ptr(g) + 2 * g_col0, X=ptr(x) + 2 * x_col0, dW=ptr(dW), rows_per_batch=rows_per_batch,
                    batches=batches, N1=N1, N2=N2, ldg=g.shape[1], ldx=x.shape[1],
                    ld1=N2 if ld1 is None else ld1, ld2=ld2, x_row_shift=x_row_shift, alpha=alpha, dbias=ptr(dbias),
                    taps=taps, tap_row_stride=tap_row_stride, tap_dw_stride=tap_dw_stride,
                    dbias_row_scale=ptr(dbias_row_scale), row_scale_div=row_scale_div)
    _call("afb_gemm_dw", C.byref(p), stream())
    return dW


def gemm_simt(a, b, M, N, K, sa, sb, *, out=None, out_dtype=torch.float32, sc=None, bias=None, alpha=1.0, beta=0.0):
    """C[i,j] = alpha*sum_k A[i*sa0 + k*sa1] B[j*sb0 + k*sb1] (+bias[j]) (+beta*C); strides in elements."""
    need_cuda(bias)
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("altformer_b200 kernels need CUDA tensors (there is no CPU fallback)")
    c = out if out is not None else torch.empty((M, N), device=a.device, dtype=out_dtype)
    sc = (N, 1) if sc is None else sc
    p = _lib.GemmSimt(A=ptr(a), B=ptr(b), C=ptr(c), bias=ptr(bias), M=M, N=N, K=K, sai=sa[0], sak=sa[1], sbj=sb[0],
                      sbk=sb[1], sci=sc[0], scj=sc[1], a_dtype=dt(a), b_dtype=dt(b), c_dtype=dt(c), alpha=alpha,
                      beta=beta)
    _call("afb_gemm_simt", C.byref(p), stream())
    return c


# ------------------------------------------------------------------------------------------------
# casts / packing
# ------------------------------------------------------------------------------------------------
def cast(x, dtype, out=None):
    need_cuda(x, out)
    y = out if out is not None else torch.empty_like(x, dtype=dtype)
    if x.numel():
        _call("afb_cast", ptr(x), dt(x), ptr(y), dt(y), x.numel(), stream())
    return y


def copy2d(src, dst, rows, cols, lds, ldd, src_off=0, dst_off=0):
    """dst[r*ldd + c + dst_off] = src[r*lds + c + src_off] with dtype conversion (offsets in elements)."""
    if not (src.is_cuda and dst.is_cuda):
        raise RuntimeError("altformer_b200 kernels need CUDA tensors (there is no CPU fallback)")
    _call("afb_copy2d", ptr(src) + src_off * src.element_size(), dt(src), lds, ptr(dst) + dst_off * dst.element_size(),
          dt(dst), ldd, rows, cols, stream())
    return dst


def cast_transpose(w):
    need_cuda(w)
    rows, cols = w.shape
    y = torch.empty((cols, rows), device=w.device, dtype=torch.bfloat16)
    _call("afb_cast_transpose", ptr(w), ptr(y), rows, cols, stream())
    return y


def conv_dw_unpack(tmp, dW, co, ci, k):
    """dW (co, ci, k) += tmp [k][co][ci] (tap-major gradient produced by gemm_dw(taps=...))."""
    need_cuda(tmp, dW)
    _call("afb_conv_dw_unpack", ptr(tmp), ptr(dW), co, ci, k, stream())
    return dW


def conv_weight_pack(w, fwd=None, bwd=None):
    """w (co, ci, k, 1) fp32 -> (fwd [co, k*ci], bwd [ci, k*co]) bf16."""
    need_cuda(w)
    co, ci, k = w.shape[:3]
    fwd = torch.empty((co, k * ci), device=w.device, dtype=torch.bfloat16) if fwd is None else fwd
    bwd = torch.empty((ci, k * co), device=w.device, dtype=torch.bfloat16) if bwd is None else bwd
    _call("afb_conv_weight_pack", ptr(w), ptr(fwd), ptr(bwd), co, ci, k, stream())
    return fwd, bwd


def split3(x, which):
    """fp32 [rows, cols] -> bf16 [rows, 3*cols]: (hi|lo|hi) for which=0, (hi|hi|lo) for which=1."""
    need_cuda(x)
    rows, cols = x.shape
    y = torch.empty((rows, 3 * cols), device=x.device, dtype=torch.bfloat16)
    _call("afb_split3", ptr(x), ptr(y), rows, cols, which, stream())
    return y


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, out_dtype=None, save_stats=True):
    need_cuda(x, gamma, beta)
    rows, D = x.shape
    y = torch.empty((rows, D), device=x.device, dtype=out_dtype or x.dtype)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    _call("afb_layernorm_fwd", ptr(x), dt(x), ptr(gamma), ptr(beta), ptr(y), dt(y), ptr(mean), ptr(rstd), rows, D,
          eps, stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=None):
    need_cuda(dy, x, gamma, mean, rstd, dgamma, dbeta, dres)
    rows, D = x.shape
    dx = torch.empty_like(x)
    _call("afb_layernorm_bwd", ptr(dy), dt(dy), ptr(x), dt(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres),
          0 if dres is None else dt(dres), ptr(dx), dt(dx), ptr(dgamma), ptr(dbeta), rows, D, stream())
    return dx


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def attention_fwd(qkv, B, L, heads, out_scale=None):
    """out_scale (B,) float32: per-sequence factor of the output (DropPath keep); factor-0 sequences are not computed."""
    need_cuda(qkv, out_scale)
    D = qkv.shape[1] // 3
    dh = D // heads
    o = torch.empty((B * L, D), device=qkv.device, dtype=qkv.dtype)
    _call("afb_attention_fwd", ptr(qkv), ptr(o), dt(qkv), B, L, heads, dh, float(dh) ** -0.5, ptr(out_scale), stream())
    return o


def attention_bwd(qkv, dO, B, L, heads):
    need_cuda(qkv, dO)
    D = qkv.shape[1] // 3
    dh = D // heads
    dqkv = torch.empty_like(qkv)
    _call("afb_attention_bwd", ptr(qkv), ptr(dO), ptr(dqkv), dt(qkv), B, L, heads, dh, float(dh) ** -0.5, stream())
    return dqkv


# ------------------------------------------------------------------------------------------------
# BatchNorm pieces
# ------------------------------------------------------------------------------------------------
def colstats(x):
    need_cuda(x)
    M, Cc = x.shape
    acc = torch.zeros((2, Cc), device=x.device, dtype=torch.float64)
    _call("afb_colstats", ptr(x), dt(x), M, Cc, Cc, ptr(acc[0]), ptr(acc[1]), stream())
    return acc


def colsum(x, out, row_scale=None, row_scale_div=1, col0=0, ncols=None):
    """out[c] += sum_m row_scale[m/div] * x[m, col0 + c] for c < ncols."""
    need_cuda(x, out, row_scale)
    M, ld = x.shape
    ncols = ld if ncols is None else ncols
    _call("afb_colsum", ptr(x) + col0 * x.element_size(), dt(x), M, ncols, ld, ptr(row_scale), row_scale_div, ptr(out),
          stream())
    return out


def bn_finalize(acc, M, gamma, beta, running_mean, running_var, momentum, eps, training):
    Cc = gamma.numel()
    out = torch.empty((4, Cc), device=gamma.device, dtype=torch.float32)  # mean, rstd, scale, shift
    _call("afb_bn_finalize", None if acc is None else ptr(acc[0]), None if acc is None else ptr(acc[1]), M, Cc,
          ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), momentum, eps, int(training), ptr(out[0]),
          ptr(out[1]), ptr(out[2]), ptr(out[3]), stream())
    return out


def bn_act_fwd(x, scale, shift, relu, res_pre=None, res_post=None, T=0, V=0, want=True, want_perm=False, out_dtype=None):
    """out_dtype: bf16 outputs from an fp32 pre-activation (the bf16 mode keeps BatchNorm inputs in fp32 so the ReLU
    mask is decided on unrounded values); default = x.dtype."""
    need_cuda(x, scale, shift, res_pre, res_post)
    M, Cc = x.shape
    od = out_dtype or x.dtype
    y = torch.empty_like(x, dtype=od) if want else None
    y2 = torch.empty_like(x, dtype=od) if want_perm else None
    res = res_pre if res_pre is not None else res_post
    _call("afb_bn_act_fwd", ptr(x), dt(x), ptr(scale), ptr(shift), ptr(res_pre), ptr(res_post),
          0 if res is None else dt(res), int(relu), ptr(y), ptr(y2), _DT[od], M, Cc, T, V, stream())
    return y, y2


def bn_bwd(dy, dy2, x, stats, gamma, beta, relu, training, dgamma, dbeta, res_pre=None, want_dres=False, T=0, V=0):
    """Two-pass BN(+ReLU) backward.  stats = (mean, rstd, scale, shift) rows from bn_finalize."""
    need_cuda(dy, dy2, x, stats, gamma, beta, dgamma, dbeta, res_pre)
    M, Cc = x.shape
    g = dy if dy is not None else dy2
    _call("afb_bn_bwd_reduce", ptr(dy), ptr(dy2), dt(g), ptr(x), dt(x), ptr(res_pre),
          0 if res_pre is None else dt(res_pre), ptr(stats[0]), ptr(stats[1]), ptr(gamma), ptr(beta), int(relu),
          ptr(dgamma), ptr(dbeta), M, Cc, T, V, stream())
    dx = torch.empty_like(x, dtype=g.dtype)      # x may be the fp32 pre-activation of the bf16 mode; gradients follow dy
    dres = torch.empty_like(x, dtype=g.dtype) if want_dres else None
    _call("afb_bn_bwd_apply", ptr(dy), ptr(dy2), dt(g), ptr(x), dt(x), ptr(res_pre),
          0 if res_pre is None else dt(res_pre), ptr(stats[0]), ptr(stats[1]), ptr(gamma), ptr(beta), ptr(dgamma),
          ptr(dbeta), int(relu), int(training), ptr(dx), ptr(dres), dt(dx), M, Cc, T, V, stream())
    return dx, dres


# ------------------------------------------------------------------------------------------------
# pooling, loss, optimizer
# ------------------------------------------------------------------------------------------------
def pool_mean_fwd(x, B, L):
    need_cuda(x)
    D = x.shape[1]
    y = torch.empty((B, D), device=x.device, dtype=x.dtype)
    _call("afb_pool_mean_fwd", ptr(x), ptr(y), dt(x), B, L, D, stream())
    return y


def pool_mean_bwd(dy, B, L):
    need_cuda(dy)
    D = dy.shape[1]
    dx = torch.empty((B * L, D), device=dy.device, dtype=dy.dtype)
    _call("afb_pool_mean_bwd", ptr(dy), ptr(dx), dt(dy), B, L, D, stream())
    return dx


def pool_max_fwd(x, B, L):
    need_cuda(x)
    D = x.shape[1]
    y = torch.empty((B, D), device=x.device, dtype=x.dtype)
    arg = torch.empty((B, D), device=x.device, dtype=torch.int32)
    _call("afb_pool_max_fwd", ptr(x), ptr(y), ptr(arg), dt(x), B, L, D, stream())
    return y, arg


def pool_max_bwd(dy, arg, B, L):
    need_cuda(dy, arg)
    D = dy.shape[1]
    dx = torch.empty((B * L, D), device=dy.device, dtype=dy.dtype)
    _call("afb_pool_max_bwd", ptr(dy), ptr(arg), ptr(dx), dt(dy), B, L, D, stream())
    return dx


def softmax_ce(logits, labels):
    need_cuda(logits, labels)
    N, Cc = logits.shape
    loss = torch.zeros((), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits)
    _call("afb_softmax_ce", ptr(logits), ptr(labels), ptr(loss), ptr(dlogits), N, Cc, stream())
    return loss, dlogits


def adamw(p, g, m, v, p_lowp, step, lr, beta1, beta2, eps, wd, grad_scale=1.0):
    need_cuda(p, g, m, v, p_lowp, step)
    _call("afb_adamw", ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_lowp), p.numel(), ptr(step), lr, beta1, beta2, eps, wd,
          grad_scale, stream())
    _call("afb_step_inc", ptr(step), stream())


def scale_rows(x, row_scale, div):
    need_cuda(x, row_scale)
    y = torch.empty_like(x)
    _call("afb_scale_rows", ptr(x), ptr(y), dt(x), x.shape[0], x.shape[1], ptr(row_scale), div, stream())
    return y


# ------------------------------------------------------------------------------------------------
# streams / ensemble
# ------------------------------------------------------------------------------------------------
def bone_stream(x, parent):
    need_cuda(x, parent)
    N, T, V, _ = x.shape
    y = torch.empty_like(x)
    _call("afb_bone_stream", ptr(x), ptr(parent), ptr(y), N * T, V, stream())
    return y


def frame_gather(x, N, T, To, V, stride, phase):
    """tokens [N*T*V, C] -> phase tensor [N*To*V, C]: frame j holds frame j*stride+phase of x (zeros past T)."""
    need_cuda(x)
    C_ = x.shape[1]
    y = torch.empty((N * To * V, C_), device=x.device, dtype=x.dtype)
    _call("afb_frame_phase", ptr(x), ptr(y), dt(x), 0, N, T, To, V, C_, stride, phase, stream())
    return y


def frame_scatter(yp, out, N, T, To, V, stride, phase):
    """inverse of frame_gather into the preallocated tokens `out` [N*T*V, C] (only the frames of this phase are written)."""
    need_cuda(yp, out)
    _call("afb_frame_phase", ptr(yp), ptr(out), dt(yp), 1, N, T, To, V, yp.shape[1], stride, phase, stream())
    return out


def mul(a, b):
    need_cuda(a, b)
    if a.dtype != b.dtype or a.numel() != b.numel():
        raise RuntimeError("mul: operands must have the same dtype and size")
    y = torch.empty_like(a)
    _call("afb_mul", ptr(a), ptr(b), ptr(y), dt(a), a.numel(), stream())
    return y


def palm_center(x, joint=1):
    need_cuda(x)
    N, T, V, _ = x.shape
    y = torch.empty_like(x)
    _call("afb_palm_center", ptr(x), ptr(y), N, T, V, joint, stream())
    return y


def augment(x, kind, params):
    """x (N, T, V, 3) fp32, kind (N,) int32, params (N, 16) fp32 -> augmented copy (afb_augment)."""
    need_cuda(x, kind, params)
    N, T, V, _ = x.shape
    if kind.dtype != torch.int32 or kind.shape != (N,) or params.dtype != torch.float32 or params.shape != (N, 16):
        raise ValueError("augment: kind must be int32 (N,), params float32 (N, 16)")
    y = torch.empty_like(x)
    _call("afb_augment", ptr(x), ptr(y), N, T, V, ptr(kind), ptr(params), stream())
    return y


def motion_stream(x):
    need_cuda(x)
    N, T, V, _ = x.shape
    y = torch.empty_like(x)
    _call("afb_motion_stream", ptr(x), ptr(y), N, T, V, stream())
    return y


def axpby(a, wa, b, wb, out=None):
    need_cuda(a, b)
    if out is None:
        out = torch.empty_like(a)
    _call("afb_axpby", ptr(a), wa, ptr(b), wb, ptr(out), a.numel(), stream())
    return out
