"""Temporal-only transformer head, drop-in for the reference's STR_TTR/TTR.py:86-221: same constructor, attribute names
and state_dict (incl. the spatial blocks the reference constructs but never runs).  forward: temporal Blocks over the T
frames of every joint -> max over frames -> Spatial_patch_to_embedding (256 -> 512, no positional term on this path) ->
mean over joints -> mlp_head (TTR.py:212-221)."""
from functools import partial

import torch
import torch.nn as nn

from altformer_b200 import functional as AF
from ..model.AltFormer.model_ST import Block, _head, prefill_drop_paths
from ..model._tokens import to_tokens


class TTR(nn.Module):
    def __init__(self, class_num, num_frame=180, num_joints=22, in_chans=128, embed_dim_ratio=256, depth=4, num_heads=8,
                 mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=None):
        super().__init__()
        self.class_num = class_num
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        embed_dim = embed_dim_ratio * 2
        self.num_frame, self.num_joints = num_frame, num_joints
        if drop_rate != 0.:
            raise ValueError("altformer_b200.TTR is built for drop_rate=0")

        self.temporal_patch_to_embedding = nn.Linear(in_chans, embed_dim_ratio)
        self.Temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frame, embed_dim_ratio))
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim_ratio))
        self.Spatial_patch_to_embedding = nn.Linear(embed_dim_ratio, embed_dim)
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, embed_dim))
        self.Spatial_cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        mk = lambda d, i: Block(dim=d, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,  # noqa: E731
                                drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
        self.Spatial_blocks = nn.ModuleList([mk(embed_dim, i) for i in range(depth)])   # constructed, never run (TTR.py:129-133)
        self.blocks = nn.ModuleList([mk(embed_dim_ratio, i) for i in range(depth)])
        self.Spatial_norm = norm_layer(embed_dim)
        self.Temporal_norm = norm_layer(embed_dim_ratio)

        self.pool = 'cls'
        self.to_latent = nn.Identity()
        self.weighted_mean = nn.Conv1d(in_channels=num_frame, out_channels=1, kernel_size=1)
        self.mlp_head = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, class_num))
        self.fcn = nn.Conv1d(512, class_num, kernel_size=1)

    def forward_tokens_nvt(self, tok, dims):
        """tok [N*V*T, in_chans] in (n, v, t) order -> logits (N, class_num) fp32."""
        N, T, V = dims
        if V != self.num_joints or T != self.num_frame:
            raise RuntimeError(f"TTR built for num_frame={self.num_frame}, num_joints={self.num_joints}; got T={T}, V={V}")
        prefill_drop_paths([(self.blocks, N * V)], tok.device)
        e = self.temporal_patch_to_embedding
        h = AF.linear(tok, e.weight, e.bias, pos=self.Temporal_pos_embed)
        for blk in self.blocks:
            h = blk.forward_rows(h, N * V, T)
        h = AF.pool_max(h, N * V, T)                            # max over frames -> [N*V, d1]   (TTR.py:166)
        e = self.Spatial_patch_to_embedding
        h = AF.linear(h, e.weight, e.bias)                      # [N*V, d2]                      (TTR.py:215)
        h = AF.pool_mean(h, N, V)                               # mean over joints -> [N, d2]    (TTR.py:217)
        return _head(self.mlp_head, h)

    def forward(self, x):
        tok, dims = to_tokens(x)
        N, T, V = dims
        tok = tok.view(N, T, V, -1).permute(0, 2, 1, 3).reshape(N * V * T, -1)   # boundary reorder for standalone use
        return self.forward_tokens_nvt(tok, dims)
