"""TCN_GCN_unit (unit_agcn branch): tcn1(gcn1(x)) + x, the layer of the "unit_agcn + temporal-conv
stack" sweep.  Reference: model/ST_TR/ST_TR_new.py:355-385 (the gcn_unit_attention branch and the
strided / channel-changing `down1` path are outside the built scope).  The residual add is fused into
the BN+ReLU pass of tcn1."""
import torch.nn as nn

from .net import Unit2D
from .unit_agcn import unit_agcn
from ._tokens import from_tokens, to_tokens


class TCN_GCN_unit(nn.Module):
    def __init__(self, in_channel, out_channel, A, kernel_size=9, stride=1, dropout=0.5, use_local_bn=False,
                 mask_learning=False, **_unused):
        super().__init__()
        if in_channel != out_channel or stride != 1:
            raise ValueError("altformer_b200.TCN_GCN_unit: only in_channel == out_channel, stride 1 is built")
        self.gcn1 = unit_agcn(in_channel, out_channel, A, use_local_bn=use_local_bn, mask_learning=mask_learning)
        self.tcn1 = Unit2D(out_channel, out_channel, kernel_size=kernel_size, dropout=dropout, stride=stride)
        self.down1 = None

    def forward_tokens(self, tok, dims):
        return self.tcn1.forward_tokens(self.gcn1.forward_tokens(tok, dims), dims, res_post=tok)

    def forward(self, x):
        tok, dims = to_tokens(x)
        return from_tokens(self.forward_tokens(tok, dims), dims)
