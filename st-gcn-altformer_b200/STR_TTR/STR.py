"""Spatial-only transformer head, drop-in for the reference's STR_TTR/STR.py:89-200: same constructor, attribute names and
state_dict (incl. the temporal blocks / embeddings / heads the reference constructs but never runs).  forward: spatial
Blocks over the V joints of every frame -> mean over joints -> linear1 (256 -> 512) -> mean over frames -> (N, 512)
features (`pred = x`, STR.py:180-191 -- the reference returns the pooled features, there is no class head on this path)."""
from functools import partial

import torch
import torch.nn as nn

from altformer_b200 import functional as AF
from ..model.AltFormer.model_ST import Block, prefill_drop_paths
from ..model._tokens import to_tokens


class STR(nn.Module):
    def __init__(self, class_num, num_frame=180, num_joints=22, in_chans=128, embed_dim_ratio=256, depth=4, num_heads=8,
                 mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=None):
        super().__init__()
        self.class_num = class_num
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        embed_dim = embed_dim_ratio * 2
        self.num_frame, self.num_joints = num_frame, num_joints
        if drop_rate != 0.:
            raise ValueError("altformer_b200.STR is built for drop_rate=0")

        self.Spatial_patch_to_embedding = nn.Linear(in_chans, embed_dim_ratio)
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, embed_dim_ratio))
        self.Spatial_cls_token = nn.Parameter(torch.randn(1, 1, embed_dim_ratio))
        self.Temporal_patch_to_embedding = nn.Linear(embed_dim_ratio, embed_dim)
        self.Temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frame, embed_dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        mk = lambda d, i: Block(dim=d, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,  # noqa: E731
                                drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
        self.Spatial_blocks = nn.ModuleList([mk(embed_dim_ratio, i) for i in range(depth)])
        self.blocks = nn.ModuleList([mk(embed_dim, i) for i in range(depth)])       # constructed, never run (STR.py:123-127)
        self.Spatial_norm = norm_layer(embed_dim_ratio)
        self.Temporal_norm = norm_layer(embed_dim)

        self.pool = 'cls'
        self.to_latent = nn.Identity()
        self.weighted_mean = nn.Conv1d(in_channels=num_frame, out_channels=1, kernel_size=1)
        self.linear1 = nn.Linear(embed_dim_ratio, embed_dim)
        self.mlp_head = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, class_num))
        self.fcn = nn.Conv1d(512, class_num, kernel_size=1)

    def forward_tokens(self, tok, dims):
        """tok [N*T*V, in_chans] in (n, t, v) order -> (N, 2*embed_dim_ratio) fp32 features."""
        N, T, V = dims
        if V != self.num_joints or T != self.num_frame:
            raise RuntimeError(f"STR built for num_frame={self.num_frame}, num_joints={self.num_joints}; got T={T}, V={V}")
        prefill_drop_paths([(self.Spatial_blocks, N * T)], tok.device)
        e = self.Spatial_patch_to_embedding
        h = AF.linear(tok, e.weight, e.bias, pos=self.Spatial_pos_embed)
        for blk in self.Spatial_blocks:
            h = blk.forward_rows(h, N * T, V)
        h = AF.pool_mean(h, N * T, V)                           # mean over joints -> [N*T, d1]   (STR.py:170)
        h = AF.linear(h, self.linear1.weight, self.linear1.bias)  # [N*T, d2]                       (STR.py:186)
        return AF.pool_mean(h, N, T).float()                    # mean over frames -> [N, d2]      (STR.py:188)

    def forward(self, x):
        tok, dims = to_tokens(x)
        return self.forward_tokens(tok, dims)
