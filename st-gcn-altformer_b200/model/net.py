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
        self.pinned_dropout_mask = None     # tests: a (rows, D_in) tensor used instead of a fresh draw
        conv_init(self.conv)

    def dropout_mask(self, tok):
        """Inverted-dropout mask as nn.Dropout applies it (net.py:40,48): Bernoulli(1 - p) / (1 - p) per element."""
        if self.pinned_dropout_mask is not None:
            return self.pinned_dropout_mask
        keep = 1.0 - self.p_drop
        return torch.empty(tok.shape, device=tok.device, dtype=torch.float32).bernoulli_(keep).div_(keep)  # RNG plumbing

    def out_dims(self, dims):
        """(N, T, V) of the output tokens: stride (s, 1) shortens the frame axis (net.py:24-27)."""
        N, T, V = dims
        return N, AF.conv_out_frames(T, self.conv.kernel_size[0], self.stride), V

    def forward_tokens(self, tok, dims, res_post=None, want_perm=False):
        if self.dim != 2:
            raise RuntimeError("altformer_b200.Unit2D: dim=3 (convolution along the joints, unused by the AltFormer / ST-GCN stacks) is not built")
        if self.p_drop > 0 and self.training:
            tok = AF.dropout(tok, self.dropout_mask(tok))
        if self.training and self.bn.track_running_stats:
            self.bn.num_batches_tracked += 1
        return AF.unit2d(tok, dims, self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias, self.bn.running_mean,
                         self.bn.running_var, self.training, self.bn.momentum, self.bn.eps, res_post, want_perm, self.stride)

    def forward(self, x):
        tok, dims = to_tokens(x)
        return from_tokens(self.forward_tokens(tok, dims), self.out_dims(dims))


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
