"""Temporal conv unit, drop-in for the reference's model/net.py (Unit2D :8-57, conv_init :60-65,
import_class :68-74).  Same constructor, attribute names (conv, bn, relu, dropout), init and
state_dict keys; forward runs one tcgen05 implicit-GEMM for the k x 1 convolution (bias fused), a
column-statistics pass and one fused BN+ReLU pass."""
import importlib
import math

import torch
import torch.nn as nn

from altformer_b200 import functional as AF
from ._tokens import from_tokens, to_tokens


class Unit2D(nn.Module):
    def __init__(self, D_in, D_out, kernel_size, stride=1, dim=2, dropout=0, bias=True):
        super().__init__()
        pad = int((kernel_size - 1) / 2)
        if dim == 2:
            self.conv = nn.Conv2d(D_in, D_out, kernel_size=(kernel_size, 1), padding=(pad, 0), stride=(stride, 1), bias=bias)
        elif dim == 3:
            self.conv = nn.Conv2d(D_in, D_out, kernel_size=(1, kernel_size), padding=(0, pad), stride=(1, stride), bias=bias)
        else:
            raise ValueError()
        self.bn = nn.BatchNorm2d(D_out)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout, inplace=False)
        self.dim, self.stride, self.p_drop = dim, stride, dropout
        conv_init(self.conv)

    def forward_tokens(self, tok, dims, res_post=None, want_perm=False):
        if self.dim != 2 or self.stride != 1:
            raise RuntimeError("altformer_b200.Unit2D: only dim=2, stride=1 (the AltFormer / agcn-stack use) is built")
        if self.p_drop > 0 and self.training:
            raise RuntimeError("altformer_b200.Unit2D: training-mode dropout > 0 is not built (AltFormer uses dropout=0)")
        if self.training and self.bn.track_running_stats:
            self.bn.num_batches_tracked += 1
        return AF.unit2d(tok, dims, self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias, self.bn.running_mean,
                         self.bn.running_var, self.training, self.bn.momentum, self.bn.eps, res_post, want_perm)

    def forward(self, x):
        tok, dims = to_tokens(x)
        return from_tokens(self.forward_tokens(tok, dims), dims)


def conv_init(module):
    n = module.out_channels
    for k in module.kernel_size:
        n = n * k
    module.weight.data.normal_(0, math.sqrt(2. / n))


def import_class(name):
    """'graph.SHRE' -> class.  Falls back to this package's own `graph` when no top-level one is importable."""
    head, *rest = name.split('.')
    try:
        mod = importlib.import_module(head)
    except ImportError:
        mod = importlib.import_module('altformer_b200.' + head)
    for comp in rest:
        mod = getattr(mod, comp)
    return mod
