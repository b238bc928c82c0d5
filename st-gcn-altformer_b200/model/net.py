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
        """(N, T, V) of the output tokens: stride (s, 1) shortens the frame axis (net.py:24-27), (1, s) the joint axis (:29-36)."""
        N, T, V = dims
        if self.dim == 3:
            return N, T, AF.conv_out_frames(V, self.conv.kernel_size[1], self.stride)
        return N, AF.conv_out_frames(T, self.conv.kernel_size[0], self.stride), V

    def _forward_tokens_dim3(self, tok, dims, res_post, want_perm):
        """dim=3 (net.py:29-36): a 1 x k convolution along the JOINTS is the k x 1 convolution of the tensor with frames and
        joints swapped, and BatchNorm statistics do not care about the order of positions -- so this variant (constructed
        nowhere in the reference's models) runs the same kernels between two token re-orderings instead of owning a second
        tap geometry."""
        if res_post is not None or want_perm:
            raise RuntimeError("altformer_b200.Unit2D(dim=3): fused residual / permuted output are only built for dim=2")
        N, T, V = dims
        C = tok.shape[1]
        swapped = tok.view(N, T, V, C).permute(0, 2, 1, 3).reshape(N * V * T, C)            # (n, v, t) order
        if self.p_drop > 0 and self.training:
            mask = self.dropout_mask(tok)
            swapped = AF.dropout(swapped, mask.view(N, T, V, C).permute(0, 2, 1, 3).reshape(N * V * T, C).contiguous())
        if self.training and self.bn.track_running_stats:
            self.bn.num_batches_tracked += 1
        y = AF.unit2d(swapped, (N, V, T), self.conv.weight.transpose(2, 3), self.conv.bias, self.bn.weight, self.bn.bias,
                      self.bn.running_mean, self.bn.running_var, self.training, self.bn.momentum, self.bn.eps, None, False, self.stride)
        Vo = self.out_dims(dims)[2]
        return y.view(N, Vo, T, -1).permute(0, 2, 1, 3).reshape(N * T * Vo, -1)

    def forward_tokens(self, tok, dims, res_post=None, want_perm=False):
        if self.dim != 2:
            return self._forward_tokens_dim3(tok, dims, res_post, want_perm)
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
