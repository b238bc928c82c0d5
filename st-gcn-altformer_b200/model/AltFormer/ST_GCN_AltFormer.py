"""Model assembly, drop-in for the reference's model/AltFormer/ST_GCN_AltFormer.py:14-87: same
constructor, attributes (gcn0, tcn0, modelA, modelB) and state_dict.  forward: the (N,T,V,3) batch is
read in place by the gcn0 kernels (no permute copy), activations stay channels-last token matrices,
tcn0 emits the (n,t,v)- and, when modelB runs, the (n,v,t)-ordered tokens in the same pass."""
import numpy as np
import torch
import torch.nn as nn

from ..net import Unit2D, import_class
from ..unit_agcn import unit_agcn
from .model_ST import ST
from .model_TS import TS


class ST_GCN_AltFormer(nn.Module):
    def __init__(self, channel, num_class, backbone_in_c=128, num_frame=180, num_joints=22, style=None, graph=None,
                 graph_args=dict(), mask_learning=False, use_local_bn=False):
        super().__init__()
        if graph is None:
            raise ValueError()
        Graph = import_class(graph) if isinstance(graph, str) else graph
        self.graph = Graph(**graph_args)
        self.A = torch.from_numpy(self.graph.A.astype(np.float32))
        self.num_frame, self.num_joints, self.num_class = num_frame, num_joints, num_class
        self.backbone_in_c, self.style = backbone_in_c, style

        self.gcn0 = unit_agcn(channel, backbone_in_c, self.A, mask_learning=mask_learning, use_local_bn=use_local_bn)
        self.tcn0 = Unit2D(backbone_in_c, backbone_in_c, kernel_size=9)
        kw = dict(num_frame=num_frame, num_joints=num_joints, in_chans=128, embed_dim_ratio=256, depth=6, num_heads=8,
                  mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_path_rate=0.1)
        self.modelA = ST(num_class, **kw)
        self.modelB = TS(num_class, **kw)

    def live_parameters(self):
        """Parameters that receive gradients for self.style (dead params and the unused stage excluded)."""
        dead = ("cls_token", "Spatial_norm", "Temporal_norm", "weighted_mean", "fcn")
        out = []
        for name, p in self.named_parameters():
            if any(d in name for d in dead):
                continue
            if self.style == 'ST' and name.startswith("modelB."):
                continue
            if self.style == 'TS' and name.startswith("modelA."):
                continue
            out.append((name, p))
        return out

    def forward(self, x):
        """x: (N, T, V, C) float tensor (CPU tensors are uploaded non-blocking) -> (N, num_class) fp32 logits."""
        dev = self.gcn0.PA.device
        if not self.gcn0.PA.is_cuda:
            raise RuntimeError("altformer_b200 runs on CUDA devices only (there is no CPU fallback); call .cuda() first")
        if not x.is_cuda:
            x = x.to(dev, non_blocking=True)
        if x.dtype != torch.float32:
            x = x.float()
        N, T, V, C = x.shape
        dims = (N, T, V)
        x = x.contiguous()
        if C == 3:
            f = self.gcn0.forward_skeleton(x)
        else:
            f = self.gcn0(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).reshape(N * T * V, -1)
        need_a, need_b = self.style != 'TS', self.style != 'ST'
        out = self.tcn0.forward_tokens(f, dims, want_perm=need_b)
        tok, tok_nvt = out if need_b else (out, None)
        if self.style == 'ST':
            return self.modelA.forward_tokens(tok, dims)
        if self.style == 'TS':
            return self.modelB.forward_tokens_nvt(tok_nvt, dims)
        x_st = self.modelA.forward_tokens(tok, dims)
        x_ts = self.modelB.forward_tokens_nvt(tok_nvt, dims)
        return AddFn.apply(x_ts, x_st)

    def load_state_dict(self, state_dict, strict=True, reference_adjacency=None, **kw):
        """Accepts the reference's DataParallel checkpoints (keys prefixed with 'module.', emsemble.py:99-104).
        reference_adjacency: None (default) = switch gcn0 to the reference's effective adjacency (A == 1e-6, see
        unit_agcn.use_reference_adjacency) exactly when the keys carry the 'module.' prefix -- i.e. the file was written
        by the reference's training scripts; True / False force it."""
        prefixed = any(k.startswith("module.") for k in state_dict)
        if prefixed:
            state_dict = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        if reference_adjacency is None:
            reference_adjacency = prefixed
        for m in self.modules():
            if hasattr(m, "use_reference_adjacency"):
                m.use_reference_adjacency(bool(reference_adjacency))
        return out


class AddFn(torch.autograd.Function):
    """pred = x_ts + x_st on device (ST_GCN_AltFormer.py:82-85)."""

    @staticmethod
    def forward(ctx, a, b):
        from altformer_b200 import ops
        return ops.axpby(a.contiguous(), 1.0, b.contiguous(), 1.0)

    @staticmethod
    def backward(ctx, g):
        return g, g
