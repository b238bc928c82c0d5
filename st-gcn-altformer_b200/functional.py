"""autograd.Function layer: each Function strings C-ABI kernel launches together for forward and
backward.  No torch arithmetic happens here; torch provides memory, streams and the autograd graph.

Precision modes (set_precision):
  'bf16' (default, performance): activations bf16, GEMM operands bf16, fp32 accumulate / statistics.
  'fp32' (parity): activations fp32; every GEMM runs on the same tcgen05 kernel as three bf16 passes
         over hi/lo-split operands (hi*hi + lo*hi + hi*lo, K-concatenated), good to ~1e-5 relative.
"""
import ctypes as C
import os
import weakref

import torch

from . import _lib, ops
from .ops import ACT_GELU, ACT_GELU_BWD, ACT_NONE

_PRECISION = ["bf16"]
_EXACT_BN = [True]   # bf16 mode: BatchNorm+ReLU pre-activations kept in fp32 and computed with hi+lo weights (see set_exact_bn_mask)
_EPOCH = [0]  # bumped by the trainer after an optimizer step that bypasses tensor version counters


def set_precision(mode):
    if mode not in ("bf16", "fp32"):
        raise ValueError(mode)
    _PRECISION[0] = mode


def get_precision():
    return _PRECISION[0]


def set_exact_bn_mask(on):
    """bf16 mode only.  A batch-stat BatchNorm followed by ReLU turns a 2^-9 rounding of its input into flipped mask
    entries (~0.1 % of them), and every flipped entry moves the gradient by its full magnitude: 2-7e-2 relative L2 on
    the layer's gradients (tools/probe_relu_mask.py reproduces it on the CPU reference).  With this switch on
    (default) the temporal conv that feeds such a BatchNorm adds the bf16 LOW half of its fp32 weights in a second
    tcgen05 pass and writes the pre-activation in fp32, so the mask is decided on unrounded values; gcn0 does the same
    inside its apply pass.  Costs one extra conv GEMM (~1 % of a cfg2 step)."""
    _EXACT_BN[0] = bool(on)


def exact_bn_mask():
    return _EXACT_BN[0] and _PRECISION[0] == "bf16"


def act_dtype():
    return torch.bfloat16 if _PRECISION[0] == "bf16" else torch.float32


def bump_weights_epoch():
    _EPOCH[0] += 1


# ------------------------------------------------------------------------------------------------
# derived (low-precision / packed) weights, cached per parameter version
# ------------------------------------------------------------------------------------------------
_cache = {}


def _slot(p):
    """Per-parameter dict of derived tensors (id-keyed: tensors cannot be WeakKeyDictionary keys)."""
    key = id(p)
    ent = _cache.get(key)
    if ent is None or ent[0]() is not p:
        ent = (weakref.ref(p, lambda _r, k=key: _cache.pop(k, None)), {})
        _cache[key] = ent
    return ent[1]


def _derived(p, kind, make):
    ver = (p._version, p.data_ptr(), _EPOCH[0])
    slot = _slot(p)
    ent = slot.get(kind)
    if ent is None or ent[0] != ver:
        with torch.no_grad():
            ent = (ver, make(p.detach()))
        slot[kind] = ent
    return ent[1]


def lowp(p):
    """bf16 copy of a parameter (the trainer installs a persistent shadow as p._afb_shadow)."""
    sh = getattr(p, "_afb_shadow", None)
    if sh is not None:
        # AdamW refreshes the shadow through raw pointers (no version bump); any other in-place edit of the parameter
        # (load_state_dict, manual init) bumps p._version and the shadow is re-cast here before it is used
        if getattr(p, "_afb_shadow_ver", None) != p._version:
            with torch.no_grad():
                ops.cast(p.detach(), torch.bfloat16, out=sh)
            p._afb_shadow_ver = p._version
        return sh
    return _derived(p, "bf16", lambda w: ops.cast(w.contiguous(), torch.bfloat16))


def _w2d(p):
    return p.detach().reshape(p.shape[0], -1)


def w_fwd(p):
    """B operand of y = x W^T: [N, K] bf16, or the (hi|hi|lo) split [N, 3K] in fp32 mode."""
    if _PRECISION[0] == "bf16":
        return lowp(p).reshape(p.shape[0], -1)
    return _derived(p, "split_b", lambda w: ops.split3(w.reshape(w.shape[0], -1).contiguous(), 1))


def w_fwd3(p):
    """(hi|hi|lo) split [N, 3K] of a weight regardless of the precision mode (exact-mask forward of the bf16 mode)."""
    return _derived(p, "split_b", lambda w: ops.split3(w.reshape(w.shape[0], -1).contiguous(), 1))


def hi_lo_cat(w32):
    """fp32 [N, K] -> bf16 [N, 2K] = (hi | lo): B operand of a two-"tap" GEMM over a bf16 A operand (tap_row_stride 0: both
    taps read the same A rows), i.e. A @ (hi + lo)^T with fp32 accumulation -- the product of bf16 activations with fp32
    weights at fp32 accuracy without materialising a split copy of the activations."""
    N, K = w32.shape
    s3 = ops.split3(w32.contiguous(), 1)                      # (hi | hi | lo)
    out = torch.empty((N, 2 * K), device=w32.device, dtype=torch.bfloat16)
    ops.copy2d(s3, out, N, K, 3 * K, 2 * K)
    ops.copy2d(s3, out, N, K, 3 * K, 2 * K, src_off=2 * K, dst_off=K)
    return out


def w_fwd2(p):
    return _derived(p, "hi_lo", lambda w: hi_lo_cat(w.reshape(w.shape[0], -1).float()))


def gemm_hi_lo(x, w2, N, **epi):
    """x [M, K] bf16 @ (hi + lo)^T -> fp32 [M, N]  (see hi_lo_cat)."""
    M, K = x.shape
    return ops.gemm_tn(x, w2, N, k_per_tap=K, taps=2, tap_row_stride=0, tap_pad=0, rows_per_batch=M, batches=1,
                       out_dtype=torch.float32, **epi)


def w_dx(p):
    """MN-major B operand of dx = dy W: W itself [N, K] (bf16), or row-stacked (hi;hi;lo) [3N, K]."""
    if _PRECISION[0] == "bf16":
        return lowp(p).reshape(p.shape[0], -1)
    return _derived(p, "split_rows", lambda w: ops.split3(w.reshape(w.shape[0], -1).contiguous(), 2).view(3 * w.shape[0], -1))


def conv_packs(p):
    """(fwd [co, k*ci], bwd [ci, k*co]) operands of the temporal conv; fp32 mode: split per tap."""
    if _PRECISION[0] == "bf16":
        return _derived(p, "conv_bf16", lambda w: ops.conv_weight_pack(w.contiguous()))

    def make(w):
        co, ci, k = w.shape[:3]
        w3 = w.reshape(co, ci, k)
        f32_fwd = w3.permute(0, 2, 1).contiguous().view(co * k, ci)             # layout plumbing on a tiny tensor
        f32_bwd = w3.flip(2).permute(1, 2, 0).contiguous().view(ci * k, co)
        return ops.split3(f32_fwd, 1).view(co, k * 3 * ci), ops.split3(f32_bwd, 1).view(ci, k * 3 * co)

    return _derived(p, "conv_split", make)


def conv_pack_lo(p):
    """forward operand [co, k*ci] of the bf16 LOW half  w - bf16(w)  of the conv weight (exact-mask second pass)."""
    def make(w):
        w = w.contiguous()
        hi32 = ops.cast(ops.cast(w, torch.bfloat16), torch.float32)
        return ops.conv_weight_pack(ops.axpby(w, 1.0, hi32, -1.0))[0]

    return _derived(p, "conv_lo", make)


def _grad_sink(p):
    """(fp32 accumulation buffer, direct?).  direct: the trainer's flat gradient view (already zeroed)."""
    gb = getattr(p, "_afb_grad", None)
    if gb is not None:
        return gb, True
    return torch.zeros(p.shape, device=p.device, dtype=torch.float32), False


def _ret(sink):
    return None if sink[1] else sink[0]


def _as_act(x):
    """Cast an activation to the current activation dtype (kernel launch; no-op if it already matches)."""
    want = act_dtype()
    if x.dtype == want:
        return x
    return ops.cast(x.contiguous(), want)


# ------------------------------------------------------------------------------------------------
# mode-aware GEMM helpers
# ------------------------------------------------------------------------------------------------
def mm_fwd(x, w_param, **epi):
    """x [M, K] (activation dtype) times w_param [N, K...]^T with the fused epilogue."""
    N = w_param.shape[0]
    if _PRECISION[0] == "bf16":
        return ops.gemm_tn(x, w_fwd(w_param), N, out_dtype=torch.bfloat16, **epi)
    return ops.gemm_tn(ops.split3(x, 0), w_fwd(w_param), N, out_dtype=torch.float32, **epi)


def mm_dx(g, w_param, **epi):
    """dx [M, K] = g [M, N] @ W [N, K]."""
    K = w_param[0].numel()
    if _PRECISION[0] == "bf16":
        return ops.gemm_tn(g, w_dx(w_param), K, b_mn_major=True, out_dtype=torch.bfloat16, **epi)
    return ops.gemm_tn(ops.split3(g, 0), w_dx(w_param), K, b_mn_major=True, out_dtype=torch.float32, **epi)


def mm_dw(g, x, dW, dbias=None, **kw):
    """dW [N, K] += g [M, N]^T @ x [M, K];  dbias [N] += column sums of g (folded into the same kernel in
    bf16 mode; a separate fp32 reduction in parity mode)."""
    if _PRECISION[0] == "bf16":
        return ops.gemm_dw(g, x, dW, dbias=dbias, **kw)
    if dbias is not None:
        ops.colsum(g, dbias)
    N1, N2 = g.shape[1], x.shape[1]
    gs, xs = ops.split3(g, 0), ops.split3(x, 1)     # (hi|lo|hi) x (hi|hi|lo)
    kw.setdefault("ld1", N2)
    kw.pop("N1", None), kw.pop("N2", None)
    for j in range(3):
        ops.gemm_dw(gs, xs, dW, N1=N1, N2=N2, g_col0=j * N1, x_col0=j * N2, **kw)
    return dW


# ------------------------------------------------------------------------------------------------
# Linear (+bias, +pos-embed, +residual, +DropPath row scale)
# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = row_scale * (x W^T + b + pos) + residual.   x [M, K]; pos [L, N] indexed by m % L."""

    @staticmethod
    def forward(ctx, x, weight, bias, pos, residual, row_scale, row_scale_div):
        x = _as_act(x)
        y = mm_fwd(x, weight, bias=bias, pos=None if pos is None else pos.detach().reshape(-1, pos.shape[-1]),
                   residual=residual, row_scale=row_scale, row_scale_div=row_scale_div)
        ctx.save_for_backward(x, weight, bias, pos, row_scale)
        ctx.div = row_scale_div
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, pos, row_scale = ctx.saved_tensors
        dy = _as_act(dy.contiguous())
        g = dy if row_scale is None else ops.scale_rows(dy, row_scale, ctx.div)
        dx = mm_dx(g, weight) if ctx.needs_input_grad[0] else None
        sw = _grad_sink(weight)
        sb = _grad_sink(bias) if bias is not None else None
        mm_dw(g, x, sw[0].view(weight.shape[0], -1), dbias=None if sb is None else sb[0])
        db = None if sb is None else _ret(sb)
        dpos = None
        if pos is not None:
            sp = _grad_sink(pos)
            L, N = pos.shape[-2], pos.shape[-1]
            ops.colsum(g.view(-1, L * N), sp[0].view(-1))
            dpos = _ret(sp)
        return dx, _ret(sw), db, dpos, (dy if ctx.has_res else None), None, None


def linear(x, weight, bias=None, pos=None, residual=None, row_scale=None, row_scale_div=1):
    return LinearFn.apply(x, weight, bias, pos, residual, row_scale, row_scale_div)


class SmallLinearFn(torch.autograd.Function):
    """Classifier head (N = num_class, not a multiple of 64): strided CUDA-core GEMM, fp32 logits."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        M, K = x.shape
        N = weight.shape[0]
        y = ops.gemm_simt(x, weight.detach(), M, N, K, (K, 1), (K, 1), bias=bias, out_dtype=torch.float32)
        ctx.save_for_backward(x, weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != torch.float32:
            dy = ops.cast(dy, torch.float32)
        M, K = x.shape
        N = weight.shape[0]
        dx = ops.gemm_simt(dy, weight.detach(), M, K, N, (N, 1), (1, K), out_dtype=x.dtype) if ctx.needs_input_grad[0] else None
        sw = _grad_sink(weight)
        ops.gemm_simt(dy, x, N, K, M, (1, N), (1, K), out=sw[0], beta=1.0)
        db = None
        if bias is not None:
            sb = _grad_sink(bias)
            one = torch.ones(1, device=dy.device, dtype=torch.float32)
            ops.gemm_simt(dy, one, N, 1, M, (1, N), (0, 0), out=sb[0].view(N, 1), sc=(1, 1), beta=1.0)
            db = _ret(sb)
        return dx, _ret(sw), db


def small_linear(x, weight, bias):
    return SmallLinearFn.apply(x, weight, bias)


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = _as_act(x)
        y, mean, rstd = ops.layernorm_fwd(x, weight.detach(), bias.detach(), eps)
        ctx.save_for_backward(x, weight, bias, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        sg, sb = _grad_sink(weight), _grad_sink(bias)
        dx = ops.layernorm_bwd(_as_act(dy.contiguous()), x, weight.detach(), mean, rstd, sg[0], sb[0])
        return dx, _ret(sg), _ret(sb), None


def layer_norm(x, weight, bias, eps):
    return LayerNormFn.apply(x, weight, bias, eps)


# ------------------------------------------------------------------------------------------------
# attention core, MLP, Block
# ------------------------------------------------------------------------------------------------
class AttentionCoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, B, L, heads):
        o = ops.attention_fwd(qkv, B, L, heads)
        ctx.save_for_backward(qkv)
        ctx.cfg = (B, L, heads)
        return o

    @staticmethod
    def backward(ctx, do):
        (qkv,) = ctx.saved_tensors
        return ops.attention_bwd(qkv, _as_act(do.contiguous()), *ctx.cfg), None, None, None


def attention_core(qkv, B, L, heads):
    return AttentionCoreFn.apply(qkv, B, L, heads)


def _fold_keep(keep, width):
    """DropPath folding (bf16 mode, width % 256 == 0): instead of materialising keep * g in the backward pass, the
    SAVED activation carries the keep factor (h_act, attention output), the following linear scales only its
    bias by it, and the backward takes dW from the pre-scaled activation, dbias from a row-scaled column sum and
    dX through the GEMM's row-scale epilogue -- no scale_rows pass over the gradient."""
    return keep is not None and _PRECISION[0] == "bf16" and width % 256 == 0


def _mlp_fwd(x, w1, b1, w2, b2, residual, keep, div):
    """returns y, h_act (carrying the keep factor when _fold_keep), h_pre"""
    if _fold_keep(keep, w2.shape[0]):
        h_act, h_pre = mm_fwd(x, w1, bias=b1, act=ACT_GELU, want_preact=True, row_scale=keep, row_scale_div=div)
        y = mm_fwd(h_act, w2, bias=b2, residual=residual, row_scale=keep, row_scale_div=div, row_scale_bias_only=True)
        return y, h_act, h_pre
    h_act, h_pre = mm_fwd(x, w1, bias=b1, act=ACT_GELU, want_preact=True)
    y = mm_fwd(h_act, w2, bias=b2, residual=residual, row_scale=keep, row_scale_div=div)
    return y, h_act, h_pre


def _mlp_bwd(g, x, h_act, h_pre, w1, b1, w2, b2, need_dx=True, keep=None, div=1):
    """g: gradient w.r.t. the fc2 output -- already DropPath-scaled unless keep is given (folded path: h_act
    carries the factor, the bias gradient and dX apply it here)."""
    s2, sb2 = _grad_sink(w2), _grad_sink(b2)
    if keep is not None:
        mm_dw(g, h_act, s2[0], dbias=sb2[0], dbias_row_scale=keep, row_scale_div=div)
        dpre = mm_dx(g, w2, act=ACT_GELU_BWD, aux=h_pre, row_scale=keep, row_scale_div=div)
    else:
        mm_dw(g, h_act, s2[0], dbias=sb2[0])
        dpre = mm_dx(g, w2, act=ACT_GELU_BWD, aux=h_pre)
    s1, sb1 = _grad_sink(w1), _grad_sink(b1)
    mm_dw(dpre, x, s1[0], dbias=sb1[0])
    dx = mm_dx(dpre, w1) if need_dx else None
    return dx, (_ret(s1), _ret(sb1), _ret(s2), _ret(sb2))


class MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x = _as_act(x)
        y, h_act, h_pre = _mlp_fwd(x, w1, b1.detach(), w2, b2.detach(), None, None, 1)
        ctx.save_for_backward(x, h_act, h_pre, w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, h_act, h_pre, w1, b1, w2, b2 = ctx.saved_tensors
        dx, grads = _mlp_bwd(_as_act(dy.contiguous()), x, h_act, h_pre, w1, b1, w2, b2, ctx.needs_input_grad[0])
        return (dx,) + grads


def mlp(x, w1, b1, w2, b2):
    return MlpFn.apply(x, w1, b1, w2, b2)


class BlockFn(torch.autograd.Function):
    """Pre-LN transformer block (model_ST.py:84-87) as one node: 4 GEMMs with fused bias / GELU /
    DropPath / residual epilogues, 2 LayerNorms, attention; backward fuses the residual-gradient adds
    into the LayerNorm backward kernels."""

    @staticmethod
    def forward(ctx, x, B, L, heads, eps, keep1, keep2, n1w, n1b, qkvw, qkvb, pw, pb, n2w, n2b, w1, b1, w2, b2):
        x = _as_act(x)
        d = lambda t: None if t is None else t.detach()  # noqa: E731
        ln1, mean1, rstd1 = ops.layernorm_fwd(x, d(n1w), d(n1b), eps)
        qkv = mm_fwd(ln1, qkvw, bias=d(qkvb))
        if _fold_keep(keep1, pw.shape[0]):   # ao carries keep1 (dropped sequences are not even computed)
            ao = ops.attention_fwd(qkv, B, L, heads, out_scale=keep1)
            x1 = mm_fwd(ao, pw, bias=d(pb), residual=x, row_scale=keep1, row_scale_div=L, row_scale_bias_only=True)
        else:
            ao = ops.attention_fwd(qkv, B, L, heads)
            x1 = mm_fwd(ao, pw, bias=d(pb), residual=x, row_scale=keep1, row_scale_div=L)
        ln2, mean2, rstd2 = ops.layernorm_fwd(x1, d(n2w), d(n2b), eps)
        x2, h_act, h_pre = _mlp_fwd(ln2, w1, d(b1), w2, d(b2), x1, keep2, L)
        ctx.save_for_backward(x, mean1, rstd1, ln1, qkv, ao, x1, mean2, rstd2, ln2, h_act, h_pre, keep1, keep2,
                              n1w, n1b, qkvw, qkvb, pw, pb, n2w, n2b, w1, b1, w2, b2)
        ctx.cfg = (B, L, heads)
        return x2

    @staticmethod
    def backward(ctx, g2):
        (x, mean1, rstd1, ln1, qkv, ao, x1, mean2, rstd2, ln2, h_act, h_pre, keep1, keep2,
         n1w, n1b, qkvw, qkvb, pw, pb, n2w, n2b, w1, b1, w2, b2) = ctx.saved_tensors
        B, L, heads = ctx.cfg
        g2 = _as_act(g2.contiguous())
        if _fold_keep(keep2, w2.shape[0]):
            dln2, (gw1, gb1, gw2, gb2) = _mlp_bwd(g2, ln2, h_act, h_pre, w1, b1, w2, b2, keep=keep2, div=L)
        else:
            gs = g2 if keep2 is None else ops.scale_rows(g2, keep2, L)
            dln2, (gw1, gb1, gw2, gb2) = _mlp_bwd(gs, ln2, h_act, h_pre, w1, b1, w2, b2)
        sg2, sb2 = _grad_sink(n2w), _grad_sink(n2b)
        g1 = ops.layernorm_bwd(dln2, x1, n2w.detach(), mean2, rstd2, sg2[0], sb2[0], dres=g2)
        sp, spb = _grad_sink(pw), _grad_sink(pb)
        if _fold_keep(keep1, pw.shape[0]):
            mm_dw(g1, ao, sp[0], dbias=spb[0], dbias_row_scale=keep1, row_scale_div=L)
            dao = mm_dx(g1, pw, row_scale=keep1, row_scale_div=L)
        else:
            gs1 = g1 if keep1 is None else ops.scale_rows(g1, keep1, L)
            mm_dw(gs1, ao, sp[0], dbias=spb[0])
            dao = mm_dx(gs1, pw)
        dqkv = ops.attention_bwd(qkv, dao, B, L, heads)
        sq = _grad_sink(qkvw)
        sqb = _grad_sink(qkvb) if qkvb is not None else None
        mm_dw(dqkv, ln1, sq[0], dbias=None if sqb is None else sqb[0])
        gqb = None if sqb is None else _ret(sqb)
        dln1 = mm_dx(dqkv, qkvw)
        sg1, sb1 = _grad_sink(n1w), _grad_sink(n1b)
        g0 = ops.layernorm_bwd(dln1, x, n1w.detach(), mean1, rstd1, sg1[0], sb1[0], dres=g1)
        return (g0, None, None, None, None, None, None, _ret(sg1), _ret(sb1), _ret(sq), gqb, _ret(sp), _ret(spb),
                _ret(sg2), _ret(sb2), gw1, gb1, gw2, gb2)


def block(x, B, L, heads, eps, keep1, keep2, *params):
    return BlockFn.apply(x, B, L, heads, eps, keep1, keep2, *params)


# ------------------------------------------------------------------------------------------------
# pooling / loss
# ------------------------------------------------------------------------------------------------
class PoolMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, B, L):
        ctx.cfg = (B, L)
        return ops.pool_mean_fwd(x, B, L)

    @staticmethod
    def backward(ctx, dy):
        return ops.pool_mean_bwd(_as_act(dy.contiguous()), *ctx.cfg), None, None


class PoolMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, B, L):
        y, arg = ops.pool_max_fwd(x, B, L)
        ctx.save_for_backward(arg)
        ctx.cfg = (B, L)
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        return ops.pool_max_bwd(_as_act(dy.contiguous()), arg, *ctx.cfg), None, None


def pool_mean(x, B, L):
    return PoolMeanFn.apply(x, B, L)


def pool_max(x, B, L):
    return PoolMaxFn.apply(x, B, L)


class CrossEntropyFn(torch.autograd.Function):
    """mean softmax cross-entropy (torch.nn.CrossEntropyLoss, train_sttran.py:161)."""

    @staticmethod
    def forward(ctx, logits, labels):
        loss, dlogits = ops.softmax_ce(logits.contiguous(), labels.contiguous())
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        # g is the scalar upstream gradient (1.0 for loss.backward()); scaled on device, no host sync
        return ops.scale_rows(dlogits, g.reshape(1), dlogits.shape[0]), None


def cross_entropy(logits, labels):
    return CrossEntropyFn.apply(logits, labels)


# ------------------------------------------------------------------------------------------------
# Unit2D: temporal conv (implicit GEMM) + BatchNorm + ReLU (+ residual after, + permuted copy)
# ------------------------------------------------------------------------------------------------
def _bn_forward(raw, bn_w, bn_b, rm, rv, training, momentum, eps):
    M = raw.shape[0]
    acc = ops.colstats(raw) if training else None
    return ops.bn_finalize(acc, M, bn_w.detach(), bn_b.detach(), rm, rv, momentum, eps, training)


def conv_out_frames(T, k, stride):
    """frames produced by Conv2d((k,1), padding ((k-1)//2, 0), stride (stride,1)) -- model/net.py:19-27"""
    return (T + 2 * ((k - 1) // 2) - k) // stride + 1


def _phase_plan(k, stride):
    """Strided k x 1 convolution as a sum of stride-1 convolutions on phase tensors: tap q = tap - pad = o*stride + phase reads
    x_phase[t' + o].  Returns [(phase, [taps in order of increasing o], o_min)] for the phases that own a tap."""
    pad = (k - 1) // 2
    plan = []
    for phase in range(stride):
        taps = [t for t in range(k) if (t - pad) % stride == phase]
        if taps:
            plan.append((phase, taps, (taps[0] - pad - phase) // stride))
    return plan


def _sub_packs(conv_w, taps):
    """(fwd, fwd_lo or None, bwd) GEMM operands of the sub-kernel made of `taps` (tiny tensors: torch indexing is layout plumbing)."""
    w = conv_w.detach()[:, :, taps, :].contiguous()
    co, ci, kk = w.shape[:3]
    if _PRECISION[0] == "bf16":
        fwd, bwd = ops.conv_weight_pack(w)
        lo = None
        if exact_bn_mask():
            hi32 = ops.cast(ops.cast(w, torch.bfloat16), torch.float32)
            lo = ops.conv_weight_pack(ops.axpby(w, 1.0, hi32, -1.0))[0]
        return fwd, lo, bwd
    w3 = w.reshape(co, ci, kk)
    f32_fwd = w3.permute(0, 2, 1).contiguous().view(co * kk, ci)
    f32_bwd = w3.flip(2).permute(1, 2, 0).contiguous().view(ci * kk, co)
    return ops.split3(f32_fwd, 1).view(co, kk * 3 * ci), None, ops.split3(f32_bwd, 1).view(ci, kk * 3 * co)


class Unit2DFn(torch.autograd.Function):
    """x tokens [N*T*V, Cin] -> relu(bn(conv_kx1(x))) (+ res_post); optionally also the (n,v,t)-ordered
    copy the TS stage consumes.  model/net.py:47-57.  stride > 1 (net.py:24-27): the output has conv_out_frames(T) frames;
    the convolution runs as one implicit GEMM per phase tensor (see _phase_plan), all accumulating into one pre-activation."""

    @staticmethod
    def forward(ctx, x, dims, conv_w, conv_b, bn_w, bn_b, rm, rv, training, momentum, eps, res_post, want_perm, stride=1):
        N, T, V = dims
        x = _as_act(x)
        co, ci, k = conv_w.shape[:3]
        three = 1 if get_precision() == "bf16" else 3
        exact = exact_bn_mask()
        bias = None if conv_b is None else conv_b.detach()
        To = conv_out_frames(T, k, stride)
        if stride == 1:
            fwd_w, _ = conv_packs(conv_w)
            a = x if three == 1 else ops.split3(x, 0)
            geo = dict(k_per_tap=three * ci, taps=k, tap_row_stride=V, tap_pad=(k - 1) // 2, rows_per_batch=T * V, batches=N)
            raw = ops.gemm_tn(a, fwd_w, co, bias=bias, out_dtype=torch.float32 if exact else act_dtype(), **geo)
            if exact:   # + x * (w - bf16(w)): the pre-activation is now the fp32 conv of the bf16 activations
                ops.gemm_tn(a, conv_pack_lo(conv_w), co, residual=raw, out=raw, out_dtype=torch.float32, **geo)
            phases = None
        else:
            raw, phases = None, []
            for phase, taps, o_min in _phase_plan(k, stride):
                xp = ops.frame_gather(x, N, T, To, V, stride, phase)
                fwd_w, lo_w, _ = _sub_packs(conv_w, taps)
                a = xp if three == 1 else ops.split3(xp, 0)
                geo = dict(k_per_tap=three * ci, taps=len(taps), tap_row_stride=V, tap_pad=-o_min, rows_per_batch=To * V, batches=N)
                if len(taps) == 1:   # a single tap needs no K blocking by taps (k_per_tap may then be any multiple of 8)
                    geo.update(taps=1, tap_row_stride=0, tap_pad=0)
                    if o_min != 0:
                        raise RuntimeError("altformer_b200.Unit2D: a single-tap phase with a frame offset is not built")
                for w_op in ((fwd_w, lo_w) if lo_w is not None else (fwd_w,)):
                    if raw is None:
                        raw = ops.gemm_tn(a, w_op, co, bias=bias, out_dtype=torch.float32 if exact else act_dtype(), **geo)
                    else:
                        ops.gemm_tn(a, w_op, co, residual=raw, out=raw, out_dtype=raw.dtype, **geo)
                phases.append(xp)
        stats = _bn_forward(raw, bn_w, bn_b, rm, rv, training, momentum, eps)
        y, y2 = ops.bn_act_fwd(raw, stats[2], stats[3], True, res_post=res_post, T=To, V=V, want_perm=want_perm,
                               out_dtype=act_dtype())
        if phases is None:
            ctx.save_for_backward(x, raw, stats, conv_w, conv_b, bn_w, bn_b)
        else:
            ctx.save_for_backward(x, raw, stats, conv_w, conv_b, bn_w, bn_b, *phases)
        ctx.cfg = (N, T, V, training, res_post is not None, want_perm, stride, To)
        if want_perm:
            return y, y2
        return y

    @staticmethod
    def backward(ctx, dy, dy2=None):
        x, raw, stats, conv_w, conv_b, bn_w, bn_b = ctx.saved_tensors[:7]
        phases = ctx.saved_tensors[7:]
        N, T, V, training, has_res, want_perm, stride, To = ctx.cfg
        co, ci, k = conv_w.shape[:3]
        dy = None if dy is None else _as_act(dy.contiguous())
        dy2 = None if dy2 is None else _as_act(dy2.contiguous())
        sg, sb = _grad_sink(bn_w), _grad_sink(bn_b)
        draw, _ = ops.bn_bwd(dy, dy2, raw, stats, bn_w.detach(), bn_b.detach(), True, training, sg[0], sb[0], T=To, V=V)
        gcb, scb = None, None
        if conv_b is not None:   # bias gradient = column sums of draw: rides on the first weight-gradient launch
            scb = _grad_sink(conv_b)
            gcb = _ret(scb)
        db = None if scb is None else scb[0]
        sw = _grad_sink(conv_w)
        flat = sw[0].view(-1)
        pad = (k - 1) // 2
        three = 1 if get_precision() == "bf16" else 3
        if stride != 1:
            # dW[:, :, tap] = sum_rows draw[row]^T x_phase[row + o V];  dx_phase = sum_taps draw[row - o V] W[:, :, tap]
            plan = _phase_plan(k, stride)
            dx = None
            if ctx.needs_input_grad[0]:   # frames of a phase no tap reads (k < stride, e.g. the 1 x 1 down1) get zero gradient
                dx = torch.empty_like(x) if len(plan) == stride else torch.zeros_like(x)
            first = True
            for (phase, taps, o_min), xp in zip(plan, phases):
                for j, tap in enumerate(taps):
                    mm_dw(draw, xp, flat[tap:], N1=co, N2=ci, rows_per_batch=To * V, batches=N, ld1=ci * k, ld2=k,
                          x_row_shift=(o_min + j) * V, dbias=db if first else None)
                    first = False
                if dx is not None:
                    _, _, bwd_w = _sub_packs(conv_w, taps)
                    g = draw if three == 1 else ops.split3(draw, 0)
                    geo = dict(k_per_tap=three * co, taps=len(taps), tap_row_stride=V, tap_pad=o_min + len(taps) - 1,
                               rows_per_batch=To * V, batches=N)
                    if len(taps) == 1:
                        geo.update(taps=1, tap_row_stride=0, tap_pad=0)
                    dxp = ops.gemm_tn(g, bwd_w, ci, out_dtype=act_dtype(), **geo)
                    ops.frame_scatter(dxp, dx, N, T, To, V, stride, phase)
            return dx, None, _ret(sw), gcb, _ret(sg), _ret(sb), None, None, None, None, None, (dy if has_res else None), None, None
        # dW[co, ci, tap] = sum_rows draw[row, co] * x[row + (tap - pad) V, ci]
        if co <= 128 and ci <= 128 and ci % 64 == 0:
            # three taps per launch share every dY tile; the tap-major scratch keeps gradient rows contiguous
            # (16-byte vector reductions) and is folded into the (co, ci, k) gradient by one small kernel
            tmp = torch.zeros((k, co, ci), device=x.device, dtype=torch.float32)
            tflat = tmp.view(-1)
            for t0 in range(0, k, 3):
                mm_dw(draw, x, tflat[t0 * co * ci:], N1=co, N2=ci, rows_per_batch=T * V, batches=N, ld1=ci, ld2=1,
                      x_row_shift=(t0 - pad) * V, taps=min(3, k - t0), tap_row_stride=V, tap_dw_stride=co * ci,
                      dbias=db if t0 == 0 else None)
            ops.conv_dw_unpack(tmp, flat, co, ci, k)
        else:
            for tap in range(k):
                mm_dw(draw, x, flat[tap:], N1=co, N2=ci, rows_per_batch=T * V, batches=N, ld1=ci * k, ld2=k,
                      x_row_shift=(tap - pad) * V, dbias=db if tap == 0 else None)
        dx = None
        if ctx.needs_input_grad[0]:
            _, bwd_w = conv_packs(conv_w)
            g = draw if three == 1 else ops.split3(draw, 0)
            dx = ops.gemm_tn(g, bwd_w, ci, k_per_tap=three * co, taps=k, tap_row_stride=V, tap_pad=pad,
                             rows_per_batch=T * V, batches=N, out_dtype=act_dtype())
        return dx, None, _ret(sw), gcb, _ret(sg), _ret(sb), None, None, None, None, None, (dy if has_res else None), None, None


def unit2d(x, dims, conv_w, conv_b, bn_w, bn_b, rm, rv, training, momentum, eps, res_post=None, want_perm=False, stride=1):
    return Unit2DFn.apply(x, dims, conv_w, conv_b, bn_w, bn_b, rm, rv, training, momentum, eps, res_post, want_perm, stride)


# ------------------------------------------------------------------------------------------------
# gcn0: unit_agcn(3 -> Cout)
# ------------------------------------------------------------------------------------------------
def _p3(ts):
    return (C.c_void_p * 3)(*[t.data_ptr() for t in ts])


_gcn0_ws = {}


def _gcn0_workspace(device):
    """Persistent zero-initialised (moments fp64 [3, 32, 96], int32[8] ticket / barrier words) pair per device; the
    kernels re-arm it ([0] / word 0: two-kernel path; [1:3] / words 4..6: fused kernel's ping-pong halves and barrier)."""
    key = (device.type, device.index)
    if key not in _gcn0_ws:
        _gcn0_ws[key] = (torch.zeros((3 * _lib.GCN0_SLOTS, _lib.GCN0_NMOM), device=device, dtype=torch.float64),
                         torch.zeros(8, device=device, dtype=torch.int32))
    return _gcn0_ws[key]


def _gcn0_struct(x, A, PA, mods, bufs, training, momentum, eps, Mmat, moments, stats, wfold, y, mma_ws=None):
    counter = _gcn0_workspace(x.device)[1]
    aop, colsum, wfrag = (None, None, None) if mma_ws is None else mma_ws
    N, T, V, _ = x.shape
    wa, ba, wb, bb, wd, bd, wdn, bdn, dng, dnb, bng, bnb = mods
    dn_rm, dn_rv, bn_rm, bn_rv = bufs
    Cout, IC = wd[0].shape[0], wa[0].shape[0]
    return _lib.Gcn0Fwd(x=x.data_ptr(), A=A.data_ptr(), PA=PA.data_ptr(), Wa=_p3(wa), ba=_p3(ba), Wb=_p3(wb), bb=_p3(bb),
                        Wd=_p3(wd), bd=_p3(bd), Wdn=wdn.data_ptr(), bdn=bdn.data_ptr(), bn_g=bng.data_ptr(),
                        bn_b=bnb.data_ptr(), dn_g=dng.data_ptr(), dn_b=dnb.data_ptr(), bn_rm=bn_rm.data_ptr(),
                        bn_rv=bn_rv.data_ptr(), dn_rm=dn_rm.data_ptr(), dn_rv=dn_rv.data_ptr(), N=N, T=T, V=V, Cout=Cout,
                        IC=IC, training=int(training), momentum=momentum, eps=eps, Mmat=Mmat.data_ptr(),
                        moments=moments.data_ptr(), counter=counter.data_ptr(), stats=stats.data_ptr(), Wfold=wfold.data_ptr(), Aop=ops.ptr(aop), colsum=ops.ptr(colsum),
                        Wfrag=ops.ptr(wfrag), y=y.data_ptr(),
                        y_dtype=ops.dt(y), precise=1 if get_precision() == "fp32" else (2 if exact_bn_mask() else 0))


class Gcn0Fn(torch.autograd.Function):
    """unit_agcn with 3 input channels on the raw (N,T,V,3) skeleton batch (model/unit_agcn.py:73-93).
    Inputs after x/A: PA, 3x(conv_a w,b), 3x(conv_b w,b), 3x(conv_d w,b), down conv w,b, down bn w,b, bn w,b,
    then the four running-stat buffers."""

    @staticmethod
    def forward(ctx, x, A, training, momentum, eps, PA, *rest):
        params, bufs = rest[:24], rest[24:]
        ops.need_cuda(x, A, PA, *params, *bufs)
        ops.ensure_device(x)
        if x.dtype != torch.float32:
            raise RuntimeError("gcn0 expects the float32 skeleton batch (N,T,V,3)")
        wa, ba = params[0:6:2], params[1:6:2]
        wb, bb = params[6:12:2], params[7:12:2]
        wd, bd = params[12:18:2], params[13:18:2]
        wdn, bdn, dng, dnb, bng, bnb = params[18:24]
        N, T, V, cin = x.shape
        if cin != 3:
            raise RuntimeError("gcn0 kernel is specialised for 3 input channels")
        Cout = wd[0].shape[0]
        dev = x.device
        Mmat = torch.empty((N, 3, V, V), device=dev, dtype=torch.float32)
        moments = _gcn0_workspace(dev)[0]
        stats = torch.empty(_lib.GCN0_NSTAT_BASE + 4 * Cout, device=dev, dtype=torch.float32)
        wfold = torch.empty((Cout, 16), device=dev, dtype=torch.float32)
        y = torch.empty((N * T * V, Cout), device=dev, dtype=act_dtype())
        det = lambda ts: [t.detach() for t in ts]  # noqa: E731
        mods = (det(wa), det(ba), det(wb), det(bb), det(wd), det(bd), wdn.detach(), bdn.detach(), dng.detach(),
                dnb.detach(), bng.detach(), bnb.detach())
        mma_ws = None
        if y.dtype == torch.bfloat16 and Cout == 128 and V <= 48:   # operands of the tensor-core apply pass
            VP = 16 if V <= 16 else (32 if V <= 32 else 48)
            mma_ws = (torch.empty(int(_lib.lib().afb_gcn0_aop_bytes(N, V)) // 2, device=dev, dtype=torch.bfloat16),
                      torch.empty((N, 3, VP), device=dev, dtype=torch.float32),
                      torch.empty((Cout // 8, 32, 2), device=dev, dtype=torch.int32))
        st = _gcn0_struct(x, A, PA.detach(), mods, bufs, training, momentum, eps, Mmat, moments, stats, wfold, y, mma_ws)
        ops._call("afb_gcn0_fwd", C.byref(st), ops.stream())
        if mma_ws is not None and V % 2 == 0 and V <= 24 and os.environ.get("AFB_GCN0_FUSED", "1")[0] != "0":
            ops.LAUNCHES[0] -= 1   # the fused cooperative kernel is one launch (the two-kernel path counts 2)
        ctx.save_for_backward(x, A, PA, Mmat, stats, wfold, y, *params, *bufs)
        ctx.cfg = (training, momentum, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        x, A, PA, Mmat, stats, wfold, y = saved[:7]
        params, bufs = saved[7:31], saved[31:]
        training, momentum, eps = ctx.cfg
        wa, ba = params[0:6:2], params[1:6:2]
        wb, bb = params[6:12:2], params[7:12:2]
        wd, bd = params[12:18:2], params[13:18:2]
        wdn, bdn, dng, dnb, bng, bnb = params[18:24]
        det = lambda ts: [t.detach() for t in ts]  # noqa: E731
        mods = (det(wa), det(ba), det(wb), det(bb), det(wd), det(bd), wdn.detach(), bdn.detach(), dng.detach(),
                dnb.detach(), bng.detach(), bnb.detach())
        Cout = wd[0].shape[0]
        moments = _gcn0_workspace(x.device)[0]  # unused by backward
        f = _gcn0_struct(x, A, PA.detach(), mods, bufs, training, momentum, eps, Mmat, moments, stats, wfold, y)
        dy = _as_act(dy.contiguous())
        ws = torch.empty(32 * Cout + 256, device=x.device, dtype=torch.float32)
        sinks = [_grad_sink(p) for p in (PA, *params)]
        gp = lambda i: sinks[i][0].data_ptr()  # noqa: E731
        b = _lib.Gcn0Bwd(f=f, dy=dy.data_ptr(), ws=ws.data_ptr(), dPA=gp(0),
                         dWa=(C.c_void_p * 3)(gp(1), gp(3), gp(5)), dba=(C.c_void_p * 3)(gp(2), gp(4), gp(6)),
                         dWb=(C.c_void_p * 3)(gp(7), gp(9), gp(11)), dbb=(C.c_void_p * 3)(gp(8), gp(10), gp(12)),
                         dWd=(C.c_void_p * 3)(gp(13), gp(15), gp(17)), dbd=(C.c_void_p * 3)(gp(14), gp(16), gp(18)),
                         dWdn=gp(19), dbdn=gp(20), ddn_g=gp(21), ddn_b=gp(22), dbn_g=gp(23), dbn_b=gp(24))
        ops._call("afb_gcn0_bwd", C.byref(b), ops.stream())
        return (None, None, None, None, None) + tuple(_ret(s) for s in sinks) + (None,) * len(bufs)


def gcn0(x, A, training, momentum, eps, PA, params, bufs):
    return Gcn0Fn.apply(x, A, training, momentum, eps, PA, *params, *bufs)


# ------------------------------------------------------------------------------------------------
# general unit_agcn(C -> Cout), C % 8 == 0  (TCN_GCN_unit stack, ST_TR_new.py:355-385)
# ------------------------------------------------------------------------------------------------
def _round_up(n, m):
    return (n + m - 1) // m * m


def _agcn_stacked(wa, ba, wb, bb, wd, bd):
    """Stacked operands, rebuilt when any member changes:
       Wab  fp32 [ldt, C]  rows (a0,a1,a2,b0,b1,b2, zero pad), bab fp32 [ldt]
       Wdc  fp32 [Cout, 3C] (Wd_0 | Wd_1 | Wd_2),             bdc fp32 [Cout] = sum_i bd_i."""
    members = (*wa, *ba, *wb, *bb, *wd, *bd)
    ver = tuple((m._version, m.data_ptr()) for m in members) + (_EPOCH[0], _PRECISION[0], exact_bn_mask())
    slot = _slot(wa[0])
    ent = slot.get("agcn_stack")
    if ent is not None and ent[0] == ver:
        return ent[1]
    with torch.no_grad():
        IC, Cin = wa[0].shape[0], wa[0].shape[1]
        Cout = wd[0].shape[0]
        ldt = _round_up(6 * IC, 64)
        dev = wa[0].device
        Wab = torch.zeros((ldt, Cin), device=dev, dtype=torch.float32)
        bab = torch.zeros(ldt, device=dev, dtype=torch.float32)
        Wdc = torch.empty((Cout, 3 * Cin), device=dev, dtype=torch.float32)
        for i in range(3):
            ops.copy2d(wa[i].detach(), Wab, IC, Cin, Cin, Cin, dst_off=i * IC * Cin)
            ops.copy2d(wb[i].detach(), Wab, IC, Cin, Cin, Cin, dst_off=(3 + i) * IC * Cin)
            ops.copy2d(ba[i].detach(), bab, 1, IC, IC, IC, dst_off=i * IC)
            ops.copy2d(bb[i].detach(), bab, 1, IC, IC, IC, dst_off=(3 + i) * IC)
            ops.copy2d(wd[i].detach(), Wdc, Cout, Cin, Cin, 3 * Cin, dst_off=i * Cin)
        bdc = ops.axpby(ops.axpby(bd[0].detach(), 1.0, bd[1].detach(), 1.0), 1.0, bd[2].detach(), 1.0)
        ab_f3 = dc_f3 = None
        if _PRECISION[0] == "bf16":
            ab_f, dc_f = ops.cast(Wab, torch.bfloat16), ops.cast(Wdc, torch.bfloat16)
            ab_x, dc_x = ab_f, dc_f
            if exact_bn_mask():   # forward operands of the exact-mask forward: (hi|lo) for the bf16 input, (hi|hi|lo) for z
                ab_f3, dc_f3 = hi_lo_cat(Wab), ops.split3(Wdc, 1)
        else:
            ab_f, dc_f = ops.split3(Wab, 1), ops.split3(Wdc, 1)
            ab_x = ops.split3(Wab, 2).view(3 * ldt, Cin)
            dc_x = ops.split3(Wdc, 2).view(3 * Cout, 3 * Cin)
        out = dict(ldt=ldt, bab=bab, bdc=bdc, ab_f=ab_f, dc_f=dc_f, ab_x=ab_x, dc_x=dc_x, ab_f3=ab_f3, dc_f3=dc_f3, Wab32=Wab)
    slot["agcn_stack"] = (ver, out)
    return out


def _gemm_raw(x, w_f, N, **epi):
    """x (activation dtype) @ pre-packed operand w_f."""
    if _PRECISION[0] == "bf16":
        return ops.gemm_tn(x, w_f, N, out_dtype=torch.bfloat16, **epi)
    return ops.gemm_tn(ops.split3(x, 0), w_f, N, out_dtype=torch.float32, **epi)


def _gemm_raw_dx(g, w_x, K, **epi):
    if _PRECISION[0] == "bf16":
        return ops.gemm_tn(g, w_x, K, b_mn_major=True, out_dtype=torch.bfloat16, **epi)
    return ops.gemm_tn(ops.split3(g, 0), w_x, K, b_mn_major=True, out_dtype=torch.float32, **epi)


def _dw_cols(g, x, dW, N1, N2, g_col0=0, x_col0=0):
    """dW [N1, N2] += g[:, g_col0:g_col0+N1]^T x[:, x_col0:x_col0+N2] (mode aware)."""
    if _PRECISION[0] == "bf16":
        return ops.gemm_dw(g, x, dW, N1=N1, N2=N2, ld1=N2, g_col0=g_col0, x_col0=x_col0)
    gs, xs = ops.split3(g, 0), ops.split3(x, 1)
    wg, wx = g.shape[1], x.shape[1]
    for j in range(3):
        ops.gemm_dw(gs, xs, dW, N1=N1, N2=N2, ld1=N2, g_col0=j * wg + g_col0, x_col0=j * wx + x_col0)
    return dW


class AgcnFn(torch.autograd.Function):
    """x tokens [N*T*V, C] -> relu(bn(sum_i conv_d_i(x M_i)) + down(x)).  Inputs after the scalars:
    PA, 3x(conv_a w,b), 3x(conv_b w,b), 3x(conv_d w,b), bn w,b, [down conv w,b, down bn w,b], then running
    stats bn_rm, bn_rv, [dn_rm, dn_rv]."""

    @staticmethod
    def forward(ctx, x, dims, A, training, momentum, eps, has_down, PA, *rest):
        N, T, V = dims
        npar = 24 if has_down else 20
        params, bufs = rest[:npar], rest[npar:]
        x = _as_act(x)
        wa, ba = params[0:6:2], params[1:6:2]
        wb, bb = params[6:12:2], params[7:12:2]
        wd, bd = params[12:18:2], params[13:18:2]
        bng, bnb = params[18:20]
        M, Cin = x.shape
        Cout, IC = wd[0].shape[0], wa[0].shape[0]
        st = _agcn_stacked(wa, ba, wb, bb, wd, bd)
        ldt = st["ldt"]
        exact = exact_bn_mask()
        # exact (bf16 mode): every value that decides a ReLU mask -- theta/phi -> M -> z -> conv_d, down -- is formed at
        # fp32 accuracy from the bf16 input (hi/lo split GEMMs, fp32 intermediates); only y and the tensors kept for the
        # backward GEMMs are bf16.  Rounding any of them to bf16 first moves the pre-activation by ~2^-9 of its typical
        # size, flips ~1e-3 of the masks and costs 3-5e-2 of gradient error against the fp32 reference.
        P = torch.empty((N, 3, V, V), device=x.device, dtype=torch.float32)
        Mmat = torch.empty_like(P)
        use_mma = (_PRECISION[0] == "bf16" and Cin % 64 == 0 and IC % 16 == 0 and (IC <= 32 or IC % 32 == 0) and V <= 48
                   and os.environ.get("AFB_AGCN_MMA", "1")[0] != "0")
        xs = None
        if exact:
            can2 = Cin % 64 == 0      # two-tap form needs whole 64-column K blocks
            x32 = None
            if can2:
                thph32 = gemm_hi_lo(x, st["ab_f3"], ldt, bias=st["bab"])
            else:
                x32 = ops.cast(x, torch.float32)
                xs = ops.split3(x32, 0)
                thph32 = ops.gemm_tn(xs, ops.split3(st["Wab32"], 1), ldt, bias=st["bab"], out_dtype=torch.float32)
            ops._call("afb_agcn_scores_fwd_mma" if use_mma else "afb_agcn_scores_fwd", ops.ptr(thph32), ops.dt(thph32), ldt, ops.ptr(A),
                      ops.ptr(PA.detach()), ops.ptr(P), ops.ptr(Mmat), N, T, V, IC, ops.stream())
            thph = ops.cast(thph32, torch.bfloat16)
            del thph32
            if use_mma:   # z leaves the tensor-core aggregate already split: (hi | lo | hi) slabs of 3C columns
                z = torch.empty((M, 9 * Cin), device=x.device, dtype=torch.bfloat16)
                ops._call("afb_agcn_aggregate_fwd_mma", ops.ptr(x), ops.ptr(Mmat), ops.ptr(z), 1, N, T, V, Cin, ops.stream())
                h_raw = ops.gemm_tn(z, st["dc_f3"], Cout, bias=st["bdc"], out_dtype=torch.float32)
            else:
                if x32 is None:
                    x32 = ops.cast(x, torch.float32)
                z32 = torch.empty((M, 3 * Cin), device=x.device, dtype=torch.float32)
                ops._call("afb_agcn_aggregate_fwd", ops.ptr(x32), ops.ptr(Mmat), ops.ptr(z32), ops.dt(x32), N, T, V, Cin, ops.stream())
                h_raw = ops.gemm_tn(ops.split3(z32, 0), st["dc_f3"], Cout, bias=st["bdc"], out_dtype=torch.float32)
                z = ops.cast(z32, torch.bfloat16)
                del z32
            del x32
        else:
            thph = _gemm_raw(x, st["ab_f"], ldt, bias=st["bab"])
            ops._call("afb_agcn_scores_fwd_mma" if use_mma else "afb_agcn_scores_fwd", ops.ptr(thph), ops.dt(thph), ldt, ops.ptr(A),
                      ops.ptr(PA.detach()), ops.ptr(P), ops.ptr(Mmat), N, T, V, IC, ops.stream())
            z = torch.empty((M, 3 * Cin), device=x.device, dtype=x.dtype)
            if use_mma:
                ops._call("afb_agcn_aggregate_fwd_mma", ops.ptr(x), ops.ptr(Mmat), ops.ptr(z), 0, N, T, V, Cin, ops.stream())
            else:
                ops._call("afb_agcn_aggregate_fwd", ops.ptr(x), ops.ptr(Mmat), ops.ptr(z), ops.dt(x), N, T, V, Cin, ops.stream())
            h_raw = _gemm_raw(z, st["dc_f"], Cout, bias=st["bdc"])
        stats_h = _bn_forward(h_raw, bng, bnb, bufs[0], bufs[1], training, momentum, eps)
        if has_down:
            wdn, bdn, dng, dnb = params[20:24]
            if exact and Cin % 64 == 0:
                d_raw = gemm_hi_lo(x, w_fwd2(wdn), Cout, bias=bdn.detach())
            elif exact:
                xs = xs if xs is not None else ops.split3(ops.cast(x, torch.float32), 0)
                d_raw = ops.gemm_tn(xs, w_fwd3(wdn), Cout, bias=bdn.detach(), out_dtype=torch.float32)
                xs = None
            else:
                d_raw = mm_fwd(x, wdn, bias=bdn.detach())
            stats_d = _bn_forward(d_raw, dng, dnb, bufs[2], bufs[3], training, momentum, eps)
            res, _ = ops.bn_act_fwd(d_raw, stats_d[2], stats_d[3], False)     # same dtype as d_raw (fp32 when exact)
        else:
            if Cin != Cout:
                raise RuntimeError("unit_agcn without `down` needs in_channels == out_channels")
            d_raw = stats_d = None
            res = x
        y, _ = ops.bn_act_fwd(h_raw, stats_h[2], stats_h[3], True, res_pre=res, out_dtype=act_dtype())
        ctx.save_for_backward(x, A, thph, P, Mmat, z, h_raw, stats_h, d_raw, stats_d, res if has_down else None, PA, *params)
        ctx.cfg = (N, T, V, training, has_down, len(bufs), use_mma)
        return y

    @staticmethod
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        x, A, thph, P, Mmat, z, h_raw, stats_h, d_raw, stats_d, res, PA = saved[:12]
        params = saved[12:]
        N, T, V, training, has_down, nbufs, use_mma = ctx.cfg
        wa, ba = params[0:6:2], params[1:6:2]
        wb, bb = params[6:12:2], params[7:12:2]
        wd, bd = params[12:18:2], params[13:18:2]
        bng, bnb = params[18:20]
        M, Cin = x.shape
        Cout, IC = wd[0].shape[0], wa[0].shape[0]
        st = _agcn_stacked(wa, ba, wb, bb, wd, bd)
        ldt = st["ldt"]
        sinks = [_grad_sink(p) for p in (PA, *params)]
        sk = lambda i: sinks[i][0]  # noqa: E731   (index 0 = PA, 1.. = params in order)
        dy = _as_act(dy.contiguous())
        dh, dres = ops.bn_bwd(dy, None, h_raw, stats_h, bng.detach(), bnb.detach(), True, training, sk(19), sk(20),
                              res_pre=res if has_down else x, want_dres=True)
        if has_down:
            wdn, bdn, dng, dnb = params[20:24]
            dd, _ = ops.bn_bwd(dres, None, d_raw, stats_d, dng.detach(), dnb.detach(), False, training, sk(23), sk(24))
            mm_dw(dd, x, sk(21).view(Cout, Cin))
            ops.colsum(dd, sk(22))
            dx = mm_dx(dd, wdn)
        else:
            dx = dres
        for i in range(3):
            _dw_cols(dh, z, sk(13 + 2 * i).view(Cout, Cin), Cout, Cin, x_col0=i * Cin)
            ops.colsum(dh, sk(14 + 2 * i))
        dz = _gemm_raw_dx(dh, st["dc_x"], 3 * Cin)
        dM = torch.empty_like(P)
        if use_mma:
            ops._call("afb_agcn_aggregate_bwd_mma", ops.ptr(x), ops.ptr(dz), ops.ptr(Mmat), ops.ptr(dx), 1, ops.ptr(dM), N, T, V, Cin,
                      ops.stream())
        else:
            ops._call("afb_agcn_aggregate_bwd", ops.ptr(x), ops.ptr(dz), ops.ptr(Mmat), ops.ptr(dx), 1, ops.ptr(dM), ops.dt(x),
                      N, T, V, Cin, ops.stream())
        dthph = torch.zeros_like(thph) if ldt != 6 * IC else torch.empty_like(thph)
        if use_mma:
            ops._call("afb_agcn_scores_bwd_mma", ops.ptr(thph), ldt, ops.ptr(P), ops.ptr(dM), ops.ptr(sk(0)), ops.ptr(dthph),
                      N, T, V, IC, ops.stream())
        else:
            ops._call("afb_agcn_scores_bwd", ops.ptr(thph), ldt, ops.ptr(P), ops.ptr(dM), ops.ptr(sk(0)), ops.ptr(dthph),
                      ops.dt(thph), N, T, V, IC, ops.stream())
        for i in range(3):
            _dw_cols(dthph, x, sk(1 + 2 * i).view(IC, Cin), IC, Cin, g_col0=i * IC)
            ops.colsum(dthph, sk(2 + 2 * i), col0=i * IC, ncols=IC)
            _dw_cols(dthph, x, sk(7 + 2 * i).view(IC, Cin), IC, Cin, g_col0=(3 + i) * IC)
            ops.colsum(dthph, sk(8 + 2 * i), col0=(3 + i) * IC, ncols=IC)
        dx = _gemm_raw_dx(dthph, st["ab_x"], Cin, residual=dx, out=dx)
        grads = tuple(_ret(s) for s in sinks)
        return (dx, None, None, None, None, None, None) + grads + (None,) * nbufs


def agcn(x, dims, A, training, momentum, eps, PA, params, bufs, has_down):
    return AgcnFn.apply(x, dims, A, training, momentum, eps, has_down, PA, *params, *bufs)


class CastFn(torch.autograd.Function):
    """Differentiable dtype change at module boundaries (afb_cast both ways)."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return ops.cast(x.contiguous(), dtype)

    @staticmethod
    def backward(ctx, g):
        return ops.cast(g.contiguous(), ctx.src), None


def to_act(x):
    """Bring a tensor to the activation dtype of the current precision mode (tracked by autograd)."""
    want = act_dtype()
    return x if x.dtype == want else CastFn.apply(x, want)


class DropoutFn(torch.autograd.Function):
    """y = x * mask with an explicit mask (values 0 or 1 / (1 - p)): nn.Dropout in training mode (model/net.py:40,48).
    The mask is a tensor so parity tests can pin it; the module draws it with torch's RNG."""

    @staticmethod
    def forward(ctx, x, mask):
        x = _as_act(x)
        mask = mask if mask.dtype == x.dtype else ops.cast(mask.contiguous(), x.dtype)
        ctx.save_for_backward(mask)
        return ops.mul(x.contiguous(), mask.contiguous())

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return ops.mul(_as_act(g.contiguous()), mask), None


def dropout(x, mask):
    return DropoutFn.apply(x, mask)


class ScaleRowsFn(torch.autograd.Function):
    """y[m, :] = s[m // div] * x[m, :]   (stand-alone DropPath)."""

    @staticmethod
    def forward(ctx, x, s, div):
        ctx.save_for_backward(s)
        ctx.div = div
        return ops.scale_rows(x.contiguous(), s, div)

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        return ops.scale_rows(g.contiguous(), s, ctx.div), None, None


def scale_rows_fn(x, s, div=1):
    return ScaleRowsFn.apply(x, s, div)
