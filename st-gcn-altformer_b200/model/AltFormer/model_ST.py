"""Spatial->Temporal AltFormer stage, drop-in for the reference's model/AltFormer/model_ST.py
(Mlp :16-32, Attention :35-67, Block :70-87, ST :89-215): same constructors, attribute names and
state_dict keys (including the never-used Spatial_cls_token, cls_token, Spatial_norm, Temporal_norm,
weighted_mean, fcn).  Tokens stay channels-last, so the reference's einops copies are views; each Block
is one autograd node of fused kernels (altformer_b200.functional.BlockFn)."""
from functools import partial

import torch
import torch.nn as nn

from altformer_b200 import functional as AF
from .._tokens import to_tokens


class DropPath(nn.Module):
    """Stochastic depth per dim-0 row (timm 0.9.12 semantics, scale_by_keep=True).  The mask is a plain
    (B,) tensor of 0 / 1/(1-p) handed to the GEMM epilogue; `pinned` lets tests supply it."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.pinned = None
        self.queue = []

    def row_scale(self, B, device):
        if self.pinned is not None:
            return self.pinned
        if self.drop_prob == 0.0 or not self.training:
            return None
        if self.queue:                      # drawn for the whole forward by prefill_drop_paths
            m = self.queue.pop(0)
            if m.shape[0] == B and m.device == device:
                return m
        keep = 1.0 - self.drop_prob
        return torch.empty(B, device=device, dtype=torch.float32).bernoulli_(keep).div_(keep)  # RNG plumbing

    def forward(self, x):
        s = self.row_scale(x.shape[0], x.device)
        if s is None:
            return x
        flat = x.reshape(x.shape[0], -1)
        return AF.scale_rows_fn(flat, s).view_as(x)


_KEEP_CACHE = {}


def prefill_drop_paths(stages, device):
    """Draw every DropPath mask of one forward (two per Block) with ONE uniform draw + compare + scale instead of a
    bernoulli_ + div_ pair per mask (40 tiny launches per cfg2 step).  `stages` = [(blocks, B), ...]; each live DropPath gets
    its two (B,) masks of 0 / 1/(1-p) queued in the order Block.forward_rows consumes them.  Same distribution as timm's
    per-call bernoulli (reference: model_ST.py:84-87 via timm DropPath); tests pin masks through `pinned` as before."""
    live = []
    for blocks, B in stages:
        for blk in blocks:
            dp = blk.drop_path
            if isinstance(dp, DropPath):
                dp.queue = []
                if dp.pinned is None and dp.drop_prob > 0.0 and dp.training:
                    live.append((dp, B))
    if not live:
        return
    key = (tuple((dp.drop_prob, B) for dp, B in live), str(device))
    vec = _KEEP_CACHE.get(key)
    if vec is None:
        keep = torch.cat([torch.full((2 * B,), 1.0 - dp.drop_prob) for dp, B in live]).to(device)
        vec = _KEEP_CACHE[key] = (keep, 1.0 / keep)
    masks = (torch.rand(vec[0].shape[0], device=device) < vec[0]) * vec[1]
    off = 0
    for dp, B in live:
        dp.queue = [masks[off:off + B], masks[off + B:off + 2 * B]]
        off += 2 * B


def _rows(x):
    """(B, L, D) -> ([B*L, D] contiguous in the activation dtype, B, L)."""
    if not x.is_cuda:
        raise RuntimeError("altformer_b200 modules run on CUDA tensors only (there is no CPU fallback)")
    B, L, D = x.shape
    return AF.to_act(x.reshape(B * L, D).contiguous()), B, L


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU or drop != 0.:
            raise ValueError("altformer_b200.Mlp is built for exact GELU and drop=0 (the AltFormer configuration)")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        shape = x.shape
        rows = AF.to_act(x.reshape(-1, shape[-1]).contiguous())
        y = AF.mlp(rows, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
        return y.view(*shape[:-1], y.shape[-1])


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if qk_scale is not None or attn_drop != 0. or proj_drop != 0.:
            raise ValueError("altformer_b200.Attention is built for qk_scale=None and zero dropout")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        rows, B, L = _rows(x)
        qkv = AF.linear(rows, self.qkv.weight, self.qkv.bias)
        o = AF.attention_core(qkv, B, L, self.num_heads)
        return AF.linear(o, self.proj.weight, self.proj.bias).view(B, L, -1)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _params(self):
        a, m = self.attn, self.mlp
        return (self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias)

    def forward_rows(self, rows, B, L):
        if self.norm1.eps != self.norm2.eps:
            raise RuntimeError("Block expects both LayerNorms to share eps")
        keep1 = keep2 = None
        if isinstance(self.drop_path, DropPath):
            keep1 = self.drop_path.row_scale(B, rows.device)
            keep2 = self.drop_path.row_scale(B, rows.device)
        return AF.block(rows, B, L, self.attn.num_heads, self.norm1.eps, keep1, keep2, *self._params())

    def forward(self, x):
        rows, B, L = _rows(x)
        return self.forward_rows(rows, B, L).view(B, L, -1)


def _head(mlp_head, rows):
    ln, fc = mlp_head[0], mlp_head[1]
    return AF.small_linear(AF.layer_norm(rows, ln.weight, ln.bias, ln.eps), fc.weight, fc.bias)


class ST(nn.Module):
    def __init__(self, class_num, num_frame=180, num_joints=22, in_chans=128, embed_dim_ratio=256, depth=4, num_heads=8,
                 mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=None):
        super().__init__()
        self.class_num = class_num
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        embed_dim = embed_dim_ratio * 2
        self.num_frame, self.num_joints = num_frame, num_joints

        self.Spatial_patch_to_embedding = nn.Linear(in_chans, embed_dim_ratio)
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, embed_dim_ratio))
        self.Spatial_cls_token = nn.Parameter(torch.randn(1, 1, embed_dim_ratio))
        self.Temporal_patch_to_embedding = nn.Linear(embed_dim_ratio, embed_dim)
        self.Temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frame, embed_dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        if drop_rate != 0.:
            raise ValueError("altformer_b200.ST is built for drop_rate=0")

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        mk = lambda d, i: Block(dim=d, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,  # noqa: E731
                                drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
        self.Spatial_blocks = nn.ModuleList([mk(embed_dim_ratio, i) for i in range(depth)])
        self.blocks = nn.ModuleList([mk(embed_dim, i) for i in range(depth)])
        self.Spatial_norm = norm_layer(embed_dim_ratio)
        self.Temporal_norm = norm_layer(embed_dim)

        self.pool = 'cls'
        self.to_latent = nn.Identity()
        self.weighted_mean = nn.Conv1d(in_channels=num_frame, out_channels=1, kernel_size=1)
        self.mlp_head = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, class_num))
        self.fcn = nn.Conv1d(512, class_num, kernel_size=1)

    def forward_tokens(self, tok, dims):
        """tok [N*T*V, in_chans] in (n, t, v) order -> logits (N, class_num) fp32."""
        N, T, V = dims
        if V != self.num_joints or T != self.num_frame:
            raise RuntimeError(f"ST built for num_frame={self.num_frame}, num_joints={self.num_joints}; got T={T}, V={V}")
        prefill_drop_paths([(self.Spatial_blocks, N * T), (self.blocks, N)], tok.device)
        e = self.Spatial_patch_to_embedding
        h = AF.linear(tok, e.weight, e.bias, pos=self.Spatial_pos_embed)
        for blk in self.Spatial_blocks:
            h = blk.forward_rows(h, N * T, V)
        h = AF.pool_mean(h, N * T, V)                       # mean over joints -> [N*T, d1]
        e = self.Temporal_patch_to_embedding
        h = AF.linear(h, e.weight, e.bias, pos=self.Temporal_pos_embed)
        for blk in self.blocks:
            h = blk.forward_rows(h, N, T)
        h = AF.pool_max(h, N, T)                            # max over frames -> [N, d2]
        return _head(self.mlp_head, h)

    def forward(self, x):
        tok, dims = to_tokens(x)
        return self.forward_tokens(tok, dims)
