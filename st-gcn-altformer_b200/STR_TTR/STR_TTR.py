"""Assembly of the ablation models, drop-in for the reference's STR_TTR/STR_TTR.py:11-84: gcn (unit_agcn 3 -> 128) ->
tcn (Unit2D 9x1) -> STR (style 'STR') or TTR (any other style); attributes gcn, tcn, modelA | modelB as in the reference."""
import numpy as np
import torch
import torch.nn as nn

from ..model.net import Unit2D, import_class
from ..model.unit_agcn import unit_agcn
from .STR import STR
from .TTR import TTR


class STR_TTR(nn.Module):
    def __init__(self, channel, num_class, backbone_in_c=128, num_frame=180, num_joints=22, style=None, graph=None,
                 graph_args=dict(), mask_learning=False, use_local_bn=False):
        super().__init__()
        if graph is None:
            raise ValueError()
        Graph = import_class(graph) if isinstance(graph, str) else graph
        self.graph = Graph(**graph_args)
        self.A = torch.from_numpy(self.graph.A.astype(np.float32))
        self.num_joints, self.num_frame, self.num_class = num_joints, num_frame, num_class
        self.backbone_in_c, self.style = backbone_in_c, style

        self.gcn = unit_agcn(channel, backbone_in_c, self.A, mask_learning=mask_learning, use_local_bn=use_local_bn)
        self.tcn = Unit2D(backbone_in_c, backbone_in_c, kernel_size=9)
        kw = dict(num_frame=num_frame, num_joints=num_joints, in_chans=128, embed_dim_ratio=256, depth=6, num_heads=8,
                  mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_path_rate=0.1)
        if style == 'STR':
            self.modelA = STR(num_class, **kw)
        else:
            self.modelB = TTR(num_class, **kw)

    def forward(self, x):
        """x: (N, T, V, C) float tensor -> STR: (N, 512) features; TTR: (N, num_class) logits (fp32)."""
        dev = self.gcn.PA.device
        if not self.gcn.PA.is_cuda:
            raise RuntimeError("altformer_b200 runs on CUDA devices only (there is no CPU fallback); call .cuda() first")
        if not x.is_cuda:
            x = x.to(dev, non_blocking=True)
        x = x.float().contiguous()
        N, T, V, C = x.shape
        dims = (N, T, V)
        if C == 3:
            f = self.gcn.forward_skeleton(x)
        else:
            f = self.gcn(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).reshape(N * T * V, -1)
        if self.style == 'STR':
            return self.modelA.forward_tokens(self.tcn.forward_tokens(f, dims, want_perm=False), dims)
        _, tok_nvt = self.tcn.forward_tokens(f, dims, want_perm=True)
        return self.modelB.forward_tokens_nvt(tok_nvt, dims)
