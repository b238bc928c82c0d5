"""CPU oracle for the ST-GCN-AltFormer hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32/fp64) functional restatement of the reference's
forward pass; gradients come from torch.autograd applied to this restatement.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.
The product path (the `st-gcn-altformer_b200` package) never does.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the *imported reference modules*
(/root/reference, with the two shims of `oracle/refshim.py`) on seeded inputs/weights and
commits the resulting vectors under `tests/golden/`; `tests/test_oracle_golden.py` checks every
function below against them (and, when /root/reference is present, against the live modules).

Reference lines each function follows (paths relative to the reference root):
  agcn_forward        model/unit_agcn.py:73-93        (+ init quirks :32-71)
  unit2d_forward      model/net.py:47-57
  tcn_gcn_forward     model/ST_TR/ST_TR_new.py:376-385 (unit_agcn branch :355-374)
  mlp_forward         model/AltFormer/model_ST.py:26-32
  attention_forward   model/AltFormer/model_ST.py:49-67
  block_forward       model/AltFormer/model_ST.py:84-87
  st_forward          model/AltFormer/model_ST.py:149-215
  ts_forward          model/AltFormer/model_TS.py:158-210
  model_forward       model/AltFormer/ST_GCN_AltFormer.py:62-87
  spatial_graph       graph/tools.py:7-22,55-60 ; graph/SHRE_graph.py:4-11 ; graph/LMDHG_graph.py:4-40
  bone / motion       data_process/Hand_Dataset.py:183-217
  ensemble            SHREC/ST_TS/emsemble.py:217-218
  train_step          SHREC/ST_TS/train_sttran.py:89-102,185-191
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# graph constants (graph/tools.py:7-22,55-60)
# --------------------------------------------------------------------------------------------
SHREC_INWARD = [(0, 2), (2, 3), (3, 4), (4, 5), (0, 1), (1, 6), (6, 7), (7, 8), (8, 9), (1, 10),
                (10, 11), (11, 12), (12, 13), (1, 14), (14, 15), (15, 16), (16, 17), (1, 18),
                (18, 19), (19, 20), (20, 21)]

_LM_ONE_HAND = [(0, 1), (1, 2), (1, 3), (1, 19), (2, 3), (2, 19), (3, 4), (4, 5), (5, 6), (3, 7),
                (7, 8), (8, 9), (9, 10), (7, 11), (11, 12), (12, 13), (13, 14), (11, 15), (15, 16),
                (16, 17), (17, 18), (15, 19), (19, 20), (20, 21), (21, 22)]
LMDHG_INWARD = _LM_ONE_HAND + [(a + 23, b + 23) for (a, b) in _LM_ONE_HAND]


def _edge_matrix(links, n):
    a = np.zeros((n, n))
    for i, j in links:
        a[j, i] = 1.0
    return a


def _col_normalise(a):
    deg = a.sum(0)
    inv = np.where(deg > 0, 1.0 / np.where(deg > 0, deg, 1.0), 0.0)
    return a @ np.diag(inv)


def spatial_graph(name: str) -> torch.Tensor:
    """(3,V,V) float32 'spatial' partition: identity, normalised inward, normalised outward."""
    if name in ("SHRE", "graph.SHRE", "shrec", 22):
        inward, n = SHREC_INWARD, 22
    elif name in ("LMDHG", "graph.LMDHG", "lmdhg", 46):
        inward, n = LMDHG_INWARD, 46
    else:
        raise ValueError(name)
    outward = [(j, i) for (i, j) in inward]
    eye = _edge_matrix([(i, i) for i in range(n)], n)
    a = np.stack([eye, _col_normalise(_edge_matrix(inward, n)), _col_normalise(_edge_matrix(outward, n))])
    return torch.from_numpy(a.astype(np.float32))


# --------------------------------------------------------------------------------------------
# state-dict layout and seeded randomisation
# --------------------------------------------------------------------------------------------
def agcn_spec(prefix, cin, cout, V):
    ic = cout // 4
    s = OrderedDict()
    s[prefix + "PA"] = (3, V, V)
    for grp, co in (("conv_a", ic), ("conv_b", ic), ("conv_d", cout)):
        for i in range(3):
            s[f"{prefix}{grp}.{i}.weight"] = (co, cin, 1, 1)
            s[f"{prefix}{grp}.{i}.bias"] = (co,)
    if cin != cout:
        s[prefix + "down.0.weight"] = (cout, cin, 1, 1)
        s[prefix + "down.0.bias"] = (cout,)
        s.update(bn_spec(prefix + "down.1.", cout))
    s.update(bn_spec(prefix + "bn.", cout))
    return s


def bn_spec(prefix, c):
    return OrderedDict([(prefix + "weight", (c,)), (prefix + "bias", (c,)), (prefix + "running_mean", (c,)),
                        (prefix + "running_var", (c,)), (prefix + "num_batches_tracked", ())])


def unit2d_spec(prefix, cin, cout, k):
    s = OrderedDict([(prefix + "conv.weight", (cout, cin, k, 1)), (prefix + "conv.bias", (cout,))])
    s.update(bn_spec(prefix + "bn.", cout))
    return s


def block_spec(prefix, d, ratio=2.0):
    h = int(d * ratio)
    return OrderedDict([
        (prefix + "norm1.weight", (d,)), (prefix + "norm1.bias", (d,)),
        (prefix + "attn.qkv.weight", (3 * d, d)), (prefix + "attn.qkv.bias", (3 * d,)),
        (prefix + "attn.proj.weight", (d, d)), (prefix + "attn.proj.bias", (d,)),
        (prefix + "norm2.weight", (d,)), (prefix + "norm2.bias", (d,)),
        (prefix + "mlp.fc1.weight", (h, d)), (prefix + "mlp.fc1.bias", (h,)),
        (prefix + "mlp.fc2.weight", (d, h)), (prefix + "mlp.fc2.bias", (d,)),
    ])


def st_spec(prefix, cls, T, V, cin=128, d1=256, depth=6):
    d2 = 2 * d1
    s = OrderedDict()
    s[prefix + "Spatial_pos_embed"] = (1, V, d1)
    s[prefix + "Spatial_cls_token"] = (1, 1, d1)
    s[prefix + "Temporal_pos_embed"] = (1, T, d2)
    s[prefix + "cls_token"] = (1, 1, d2)
    s[prefix + "Spatial_patch_to_embedding.weight"] = (d1, cin)
    s[prefix + "Spatial_patch_to_embedding.bias"] = (d1,)
    s[prefix + "Temporal_patch_to_embedding.weight"] = (d2, d1)
    s[prefix + "Temporal_patch_to_embedding.bias"] = (d2,)
    for i in range(depth):
        s.update(block_spec(f"{prefix}Spatial_blocks.{i}.", d1))
    for i in range(depth):
        s.update(block_spec(f"{prefix}blocks.{i}.", d2))
    s[prefix + "Spatial_norm.weight"] = (d1,)
    s[prefix + "Spatial_norm.bias"] = (d1,)
    s[prefix + "Temporal_norm.weight"] = (d2,)
    s[prefix + "Temporal_norm.bias"] = (d2,)
    s[prefix + "weighted_mean.weight"] = (1, T, 1)
    s[prefix + "weighted_mean.bias"] = (1,)
    s[prefix + "mlp_head.0.weight"] = (d2,)
    s[prefix + "mlp_head.0.bias"] = (d2,)
    s[prefix + "mlp_head.1.weight"] = (cls, d2)
    s[prefix + "mlp_head.1.bias"] = (cls,)
    s[prefix + "fcn.weight"] = (cls, 512, 1)
    s[prefix + "fcn.bias"] = (cls,)
    return s


def ts_spec(prefix, cls, T, V, cin=128, d1=256, depth=6):
    d2 = 2 * d1
    s = OrderedDict()
    s[prefix + "Temporal_pos_embed"] = (1, T, d1)
    s[prefix + "cls_token"] = (1, 1, d1)
    s[prefix + "Spatial_pos_embed"] = (1, V, d2)
    s[prefix + "Spatial_cls_token"] = (1, 1, d2)
    s[prefix + "temporal_patch_to_embedding.weight"] = (d1, cin)
    s[prefix + "temporal_patch_to_embedding.bias"] = (d1,)
    s[prefix + "Spatial_patch_to_embedding.weight"] = (d2, d1)
    s[prefix + "Spatial_patch_to_embedding.bias"] = (d2,)
    for i in range(depth):
        s.update(block_spec(f"{prefix}Spatial_blocks.{i}.", d2))
    for i in range(depth):
        s.update(block_spec(f"{prefix}blocks.{i}.", d1))
    s[prefix + "Spatial_norm.weight"] = (d2,)
    s[prefix + "Spatial_norm.bias"] = (d2,)
    s[prefix + "Temporal_norm.weight"] = (d1,)
    s[prefix + "Temporal_norm.bias"] = (d1,)
    s[prefix + "weighted_mean.weight"] = (1, T, 1)
    s[prefix + "weighted_mean.bias"] = (1,)
    s[prefix + "mlp_head.0.weight"] = (d2,)
    s[prefix + "mlp_head.0.bias"] = (d2,)
    s[prefix + "mlp_head.1.weight"] = (cls, d2)
    s[prefix + "mlp_head.1.bias"] = (cls,)
    s[prefix + "fcn.weight"] = (cls, 512, 1)
    s[prefix + "fcn.bias"] = (cls,)
    return s


def model_spec(channel=3, cls=28, T=32, V=22, backbone=128):
    s = OrderedDict()
    s.update(agcn_spec("gcn0.", channel, backbone, V))
    s.update(unit2d_spec("tcn0.", backbone, backbone, 9))
    s.update(st_spec("modelA.", cls, T, V, backbone))
    s.update(ts_spec("modelB.", cls, T, V, backbone))
    return s


def random_state(spec, seed=0, dtype=torch.float32):
    """Seeded non-degenerate values for every entry of `spec` (the reference's own
    init makes the GCN branch invisible, so parity on it would be vacuous)."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape in spec.items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.int64)
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("PA"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif "pos_embed" in name or "cls_token" in name:
            t = 0.02 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 1:  # BN / LN gamma
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "bias":
            is_norm = any(k in name for k in (".bn.", "down.1.", "norm", "mlp_head.0."))
            t = (0.1 if is_norm else 0.02) * torch.randn(shape, generator=g)
        else:  # conv / linear weights: fan-in scaled
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            t = torch.randn(shape, generator=g) / math.sqrt(max(fan_in, 1))
        out[name] = t.to(dtype) if t.is_floating_point() else t
    return out


def synthetic_batch(N, T, V, cls, seed=1234):
    """BASELINE.md §4 inputs: 0.2*randn skeletons, palm-centred (Hand_Dataset.py:61), labels."""
    g = torch.Generator().manual_seed(seed)
    x = 0.2 * torch.randn(N, T, V, 3, generator=g)
    x = x - x[:, :1, 1:2, :]
    y = torch.randint(0, cls, (N,), generator=g)
    return x, y


# --------------------------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------------------------
def _bn(x, p, prefix, training, eps=1e-5, momentum=0.1):
    rm, rv = p.get(prefix + "running_mean"), p.get(prefix + "running_var")
    return F.batch_norm(x, rm, rv, p[prefix + "weight"], p[prefix + "bias"], training, momentum, eps)


def agcn_attention(x, p, prefix, A):
    """(N,3,V,V) mixing matrices M_i = softmax_u(theta_i^T phi_i / (IC*T)) + A_i + PA_i."""
    N, C, T, V = x.shape
    aeff = A.to(x.dtype) + p[prefix + "PA"]
    mats = []
    for i in range(3):
        th = F.conv2d(x, p[f"{prefix}conv_a.{i}.weight"], p[f"{prefix}conv_a.{i}.bias"])
        ph = F.conv2d(x, p[f"{prefix}conv_b.{i}.weight"], p[f"{prefix}conv_b.{i}.bias"])
        ic = th.shape[1]
        a1 = th.permute(0, 3, 1, 2).reshape(N, V, ic * T)
        a2 = ph.reshape(N, ic * T, V)
        s = torch.matmul(a1, a2) / (ic * T)
        mats.append(torch.softmax(s, dim=-2) + aeff[i])
    return torch.stack(mats, 1)


def agcn_forward(x, p, prefix, A, training=False):
    """x (N,C,T,V) -> (N,C_out,T,V).  model/unit_agcn.py:73-93."""
    N, C, T, V = x.shape
    M = agcn_attention(x, p, prefix, A)
    h = None
    for i in range(3):
        z = torch.matmul(x.reshape(N, C * T, V), M[:, i]).view(N, C, T, V)
        hi = F.conv2d(z, p[f"{prefix}conv_d.{i}.weight"], p[f"{prefix}conv_d.{i}.bias"])
        h = hi if h is None else h + hi
    y = _bn(h, p, prefix + "bn.", training)
    if prefix + "down.0.weight" in p:
        d = _bn(F.conv2d(x, p[prefix + "down.0.weight"], p[prefix + "down.0.bias"]), p, prefix + "down.1.", training)
    else:
        d = x
    return torch.relu(y + d)


def unit2d_forward(x, p, prefix, training=False, stride=1):
    """Conv(k x 1, pad (k-1)//2) -> BN -> ReLU; dropout p=0.  model/net.py:47-57.  A (1, k) weight is the dim=3 variant
    (convolution along the joints, net.py:29-36)."""
    w = p[prefix + "conv.weight"]
    if w.shape[2] == 1 and w.shape[3] > 1:
        k = w.shape[3]
        y = F.conv2d(x, w, p.get(prefix + "conv.bias"), stride=(1, stride), padding=(0, (k - 1) // 2))
    else:
        k = w.shape[2]
        y = F.conv2d(x, w, p.get(prefix + "conv.bias"), stride=(stride, 1), padding=((k - 1) // 2, 0))
    return torch.relu(_bn(y, p, prefix + "bn.", training))


def boundary_bf16(x):
    """Module-boundary quantiser for checking the bf16 mode: the SAME fp32 reference math, but the activation handed from
    one conv+BN+ReLU module to the next is the bf16 value the CUDA path stores there (straight-through gradient).  Without
    it a composite of two such modules cannot be compared tighter than ~4e-2 on gradients: the second module's ReLU masks
    are decided on inputs that differ by the storage rounding (2^-9), which flips ~1e-3 of them."""
    return x + (x.detach().bfloat16().to(x.dtype) - x.detach())


def tcn_gcn_forward(x, p, prefix, A, training=False, boundary=None, stride=1):
    """tcn1(gcn1(x)) + (x | down1(x)).  ST_TR_new.py:376-385; down1 = Unit2D(k=1, stride) exists when the channel count or
    the stride changes (:369-374) -- present here iff its weights are in `p`."""
    h = agcn_forward(x, p, prefix + "gcn1.", A, training)
    if boundary is not None:
        h = boundary(h)
    res = unit2d_forward(x, p, prefix + "down1.", training, stride) if prefix + "down1.conv.weight" in p else x
    return unit2d_forward(h, p, prefix + "tcn1.", training, stride) + res


def mlp_forward(x, p, prefix):
    h = F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"])
    return F.linear(F.gelu(h), p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def attention_forward(x, p, prefix, heads=8):
    B, L, D = x.shape
    dh = D // heads
    qkv = F.linear(x, p[prefix + "qkv.weight"], p.get(prefix + "qkv.bias"))
    qkv = qkv.view(B, L, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * dh ** -0.5, dim=-1)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, L, D)
    return F.linear(o, p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def block_forward(x, p, prefix, heads=8, keep=None):
    """Pre-LN residual block, LN eps 1e-6.  `keep` = optional (B,) DropPath scale (0 or 1/(1-p))
    applied to both branches with *independent* masks: keep = (mask_attn, mask_mlp)."""
    D = x.shape[-1]
    a = attention_forward(F.layer_norm(x, (D,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-6),
                          p, prefix + "attn.", heads)
    if keep is not None:
        a = a * keep[0].view(-1, 1, 1)
    x = x + a
    m = mlp_forward(F.layer_norm(x, (D,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-6), p, prefix + "mlp.")
    if keep is not None:
        m = m * keep[1].view(-1, 1, 1)
    return x + m


def _stage(x, p, embed, pos, blocks_prefix, depth, heads, keeps):
    x = F.linear(x, p[embed + ".weight"], p[embed + ".bias"]) + p[pos]
    for i in range(depth):
        x = block_forward(x, p, f"{blocks_prefix}.{i}.", heads, None if keeps is None else keeps[i])
    return x


def _head(x, p, prefix):
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), p[prefix + "mlp_head.0.weight"], p[prefix + "mlp_head.0.bias"], 1e-5)
    return F.linear(x, p[prefix + "mlp_head.1.weight"], p[prefix + "mlp_head.1.bias"])


def _depth(p, prefix):
    d = 0
    while f"{prefix}.{d}.norm1.weight" in p:
        d += 1
    return d


def st_forward(x, p, prefix, heads=8, keeps=None):
    """x (N,C,T,V) -> logits.  keeps = (spatial_keeps, temporal_keeps) per block or None."""
    N, C, T, V = x.shape
    tok = x.permute(0, 2, 3, 1).reshape(N * T, V, C)
    h = _stage(tok, p, prefix + "Spatial_patch_to_embedding", prefix + "Spatial_pos_embed",
               prefix + "Spatial_blocks", _depth(p, prefix + "Spatial_blocks"), heads, None if keeps is None else keeps[0])
    h = h.mean(dim=1).view(N, T, -1)
    h = _stage(h, p, prefix + "Temporal_patch_to_embedding", prefix + "Temporal_pos_embed",
               prefix + "blocks", _depth(p, prefix + "blocks"), heads, None if keeps is None else keeps[1])
    return _head(h.max(dim=1).values, p, prefix)


def ts_forward(x, p, prefix, heads=8, keeps=None):
    N, C, T, V = x.shape
    tok = x.permute(0, 3, 2, 1).reshape(N * V, T, C)
    h = _stage(tok, p, prefix + "temporal_patch_to_embedding", prefix + "Temporal_pos_embed",
               prefix + "blocks", _depth(p, prefix + "blocks"), heads, None if keeps is None else keeps[0])
    h = h.max(dim=1).values.view(N, V, -1)
    h = _stage(h, p, prefix + "Spatial_patch_to_embedding", prefix + "Spatial_pos_embed",
               prefix + "Spatial_blocks", _depth(p, prefix + "Spatial_blocks"), heads, None if keeps is None else keeps[1])
    return _head(h.mean(dim=1), p, prefix)


def str_spec(prefix, cls, T, V, cin=128, d1=256, depth=6):
    """state_dict of the reference's STR (STR_TTR/STR.py:89-148): ST's entries plus linear1."""
    s = st_spec(prefix, cls, T, V, cin, d1, depth)
    s[prefix + "linear1.weight"] = (2 * d1, d1)
    s[prefix + "linear1.bias"] = (2 * d1,)
    return s


def strttr_spec(style, channel=3, cls=28, T=32, V=22, backbone=128):
    """STR_TTR assembly (STR_TTR/STR_TTR.py:11-60): gcn, tcn and modelA = STR (style 'STR') or modelB = TTR (same entries as TS)."""
    s = OrderedDict()
    s.update(agcn_spec("gcn.", channel, backbone, V))
    s.update(unit2d_spec("tcn.", backbone, backbone, 9))
    s.update(str_spec("modelA.", cls, T, V, backbone) if style == "STR" else ts_spec("modelB.", cls, T, V, backbone))
    return s


def str_forward(x, p, prefix, heads=8, keeps=None):
    """STR.forward (STR_TTR/STR.py:150-191): spatial stage -> mean over joints -> linear1 -> mean over frames -> (N, 2 d1)."""
    N, C, T, V = x.shape
    tok = x.permute(0, 2, 3, 1).reshape(N * T, V, C)
    h = _stage(tok, p, prefix + "Spatial_patch_to_embedding", prefix + "Spatial_pos_embed",
               prefix + "Spatial_blocks", _depth(p, prefix + "Spatial_blocks"), heads, keeps)
    h = h.mean(dim=1).view(N, T, -1)
    h = F.linear(h, p[prefix + "linear1.weight"], p[prefix + "linear1.bias"])
    return h.mean(dim=1)


def ttr_forward(x, p, prefix, heads=8, keeps=None):
    """TTR.forward (STR_TTR/TTR.py:151-221): temporal stage -> max over frames -> Spatial_patch_to_embedding (no positional
    term) -> mean over joints -> mlp_head."""
    N, C, T, V = x.shape
    tok = x.permute(0, 3, 2, 1).reshape(N * V, T, C)
    h = _stage(tok, p, prefix + "temporal_patch_to_embedding", prefix + "Temporal_pos_embed",
               prefix + "blocks", _depth(p, prefix + "blocks"), heads, keeps)
    h = h.max(dim=1).values.view(N, V, -1)
    h = F.linear(h, p[prefix + "Spatial_patch_to_embedding.weight"], p[prefix + "Spatial_patch_to_embedding.bias"])
    return _head(h.mean(dim=1), p, prefix)


def strttr_forward(x, p, A, style="STR", training=False, keeps=None):
    """STR_TTR.forward (STR_TTR/STR_TTR.py:62-84): x (N,T,V,3) -> STR features (N, 512) | TTR logits (N, cls)."""
    x = x.permute(0, 3, 1, 2).contiguous()
    f = unit2d_forward(agcn_forward(x, p, "gcn.", A, training), p, "tcn.", training)
    return str_forward(f, p, "modelA.", keeps=keeps) if style == "STR" else ttr_forward(f, p, "modelB.", keeps=keeps)


def backbone_forward(x, p, A, training=False, boundary=None):
    """(N,T,V,3) -> (N,128,T,V): gcn0 then tcn0."""
    x = x.permute(0, 3, 1, 2).contiguous()
    h = agcn_forward(x, p, "gcn0.", A, training)
    if boundary is not None:
        h = boundary(h)
    return unit2d_forward(h, p, "tcn0.", training)


def model_forward(x, p, A, style="ST", training=False, keeps=None, boundary=None):
    """ST_GCN_AltFormer.forward: x (N,T,V,3) -> (N,cls).  keeps: dict {'A':..., 'B':...} or None.
    boundary: optional module-boundary quantiser (boundary_bf16) applied to the gcn0 output."""
    f = backbone_forward(x, p, A, training, boundary)
    ka = None if keeps is None else keeps.get("A")
    kb = None if keeps is None else keeps.get("B")
    if style == "ST":
        return st_forward(f, p, "modelA.", keeps=ka)
    if style == "TS":
        return ts_forward(f, p, "modelB.", keeps=kb)
    return ts_forward(f, p, "modelB.", keeps=kb) + st_forward(f, p, "modelA.", keeps=ka)


# --------------------------------------------------------------------------------------------
# streams / ensemble (SURVEY §8f "next" rows)
# --------------------------------------------------------------------------------------------
BONE_START = [0, 0, 0, 2, 3, 4, 1, 6, 7, 8, 1, 10, 11, 12, 1, 14, 15, 16, 1, 18, 19, 20]


def bone_stream(x):
    """x (N,T,22,3): joint minus its parent (Hand_Dataset.py:200-217)."""
    return x - x[:, :, BONE_START, :]


def motion_stream(x):
    """x (N,T,V,3): next frame minus this frame, last frame zero (Hand_Dataset.py:183-198)."""
    out = torch.zeros_like(x)
    out[:, :-1] = x[:, 1:] - x[:, :-1]
    return out


def augment(x, kind, params):
    """Hand_Dataset.data_aug with the random draws made explicit (data_process/Hand_Dataset.py:84-157).  x (N, T, V, 3);
    kind (N,) ints: 0 scale (:86-96) | 1 shift (:98-107) | 2 noise on four joints (:109-123) | 3 time_interpolate (:125-142);
    params (N, 16): factor | offset xyz | 4 joint ids + 4 x xyz offsets | r."""
    out = x.clone()
    N, T, V, _ = x.shape
    for n in range(N):
        k, p = int(kind[n]), params[n]
        if k == 0:
            out[n] = x[n] * p[0]
        elif k == 1:
            out[n] = x[n] + p[:3]
        elif k == 2:
            for q in range(4):
                out[n, :, int(p[q])] += p[4 + 3 * q: 7 + 3 * q]
        elif k == 3 and T > 1:
            res = x[n, :-1] + p[0] * (x[n, 1:] - x[n, :-1])          # T - 1 interpolated frames ...
            out[n] = torch.cat([res, res[-1:]], 0)                    # ... padded with the last one to time_len
    return out


def ensemble(logits_st, logits_ts):
    return 0.8 * logits_st + 0.2 * logits_ts


# --------------------------------------------------------------------------------------------
# training step (train_sttran.py:89-102,185-191): fwd -> CE -> zero_grad -> bwd -> AdamW
# --------------------------------------------------------------------------------------------
def adamw_step(param, grad, m, v, step, lr=2e-4, b1=0.9, b2=0.999, eps=1e-8, wd=0.1):
    """torch.optim.AdamW semantics (decoupled decay, bias-corrected), in place on clones."""
    param = param * (1 - lr * wd)
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    mhat = m / (1 - b1 ** step)
    denom = (v.sqrt() / math.sqrt(1 - b2 ** step)) + eps
    return param - lr * mhat / denom, m, v


class OracleModel(torch.nn.Module):
    """nn.Module wrapper so the oracle can be timed as the CPU baseline with a stock optimizer."""

    def __init__(self, state, A, style="ST"):
        super().__init__()
        self.names = list(state.keys())
        self.A = A
        self.style = style
        self._p = torch.nn.ParameterDict()
        self._b = {}
        for k, t in state.items():
            if t.is_floating_point() and "running_" not in k:
                self._p[k.replace(".", "/")] = torch.nn.Parameter(t.clone())
            else:
                self._b[k] = t.clone()

    def tensors(self):
        d = {k.replace("/", "."): v for k, v in self._p.items()}
        d.update(self._b)
        return d

    def forward(self, x):
        return model_forward(x, self.tensors(), self.A, self.style, self.training)
