"""Adaptive graph convolution, drop-in for the reference's model/unit_agcn.py:31-93.

Same constructor signature, attribute names (PA, conv_a, conv_b, conv_d, down, bn, soft, relu) and
init (conv_init / bn_init / conv_branch_init, :12-28,64-71), hence the same state_dict keys.  Two
deliberate, documented deltas: `A` is *copied* into a non-persistent buffer (the
reference aliases it with PA and overwrites it with 1e-6; `use_reference_adjacency()` reproduces that
for checkpoints the reference trained), and `A` lives on the module's device instead of being
re-uploaded every forward (:75).

forward: C_in == 3 -> the fused gcn0 kernels (Gram-form scores, moment-trick BN, one write pass);
otherwise the general path (tcgen05 GEMMs for theta/phi, conv_d, down + per-sample graph kernels).
"""
import math

import torch
import torch.nn as nn

from altformer_b200 import functional as AF
from ._tokens import from_tokens, to_tokens


def conv_init(conv):
    nn.init.kaiming_normal_(conv.weight, mode='fan_out')
    nn.init.constant_(conv.bias, 0)


def bn_init(bn, scale):
    nn.init.constant_(bn.weight, scale)
    nn.init.constant_(bn.bias, 0)


def conv_branch_init(conv, branches):
    n, k1, k2 = conv.weight.shape[:3]
    nn.init.normal_(conv.weight, 0, math.sqrt(2. / (n * k1 * k2 * branches)))
    nn.init.constant_(conv.bias, 0)


class unit_agcn(nn.Module):
    def __init__(self, in_channels, out_channels, A, coff_embedding=4, num_subset=3, use_local_bn=False, mask_learning=False):
        super().__init__()
        if num_subset != 3:
            raise ValueError("altformer_b200.unit_agcn is built for the 3-subset spatial partition")
        self.inter_c = out_channels // coff_embedding
        self.in_channels, self.out_channels, self.num_subset = in_channels, out_channels, num_subset
        A = torch.as_tensor(A, dtype=torch.float32)
        self.PA = nn.Parameter(torch.full_like(A, 1e-6))
        self.register_buffer("A", A.clone(), persistent=False)
        self.register_buffer("A_graph", A.clone(), persistent=False)   # the normalised adjacency, kept for use_reference_adjacency(False)
        self.conv_a, self.conv_b, self.conv_d = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _ in range(num_subset):
            self.conv_a.append(nn.Conv2d(in_channels, self.inter_c, 1))
            self.conv_b.append(nn.Conv2d(in_channels, self.inter_c, 1))
            self.conv_d.append(nn.Conv2d(in_channels, out_channels, 1))
        if in_channels != out_channels:
            self.down = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        else:
            self.down = lambda x: x
        self.bn = nn.BatchNorm2d(out_channels)
        self.soft = nn.Softmax(-2)
        self.relu = nn.ReLU()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                conv_init(m)
            elif isinstance(m, nn.BatchNorm2d):
                bn_init(m, 1)
        bn_init(self.bn, 1e-6)
        for i in range(num_subset):
            conv_branch_init(self.conv_d[i], num_subset)

    def use_reference_adjacency(self, on=True):
        """Reference-compat mode for weights TRAINED BY THE REFERENCE.  The reference builds `PA = nn.Parameter(A)` and then
        `constant_(PA, 1e-6)` (model/unit_agcn.py:36-38): the parameter aliases the caller's tensor, so `self.A` (:38) is
        overwritten with 1e-6 too, and after `model.cuda()` it stays a CPU constant 1e-6 that forward() adds to PA (:75-76).
        A checkpoint saved by the reference's scripts therefore holds a PA that was learnt against A == 1e-6, not against
        the normalised graph adjacency this module is constructed with.  on=True fills the `A` buffer with 1e-6 so such a
        checkpoint reproduces the reference's logits; on=False restores the graph adjacency."""
        with torch.no_grad():
            if on:
                self.A.fill_(1e-6)
            else:
                self.A.copy_(self.A_graph)
        return self

    @property
    def has_down(self):
        return isinstance(self.down, nn.Module)

    def _abd(self):
        out = []
        for group in (self.conv_a, self.conv_b, self.conv_d):
            for conv in group:
                out += [conv.weight, conv.bias]
        return out

    def _bump(self):
        if self.training:
            self.bn.num_batches_tracked += 1
            if self.has_down:
                self.down[1].num_batches_tracked += 1

    def forward_skeleton(self, x):
        """x: the raw (N, T, V, 3) float32 batch -> tokens [N*T*V, C_out] (gcn0 fast path)."""
        if not self.has_down:
            raise RuntimeError("gcn0 path needs in_channels (3) != out_channels")
        self._bump()
        dn_conv, dn_bn = self.down[0], self.down[1]
        params = self._abd() + [dn_conv.weight, dn_conv.bias, dn_bn.weight, dn_bn.bias, self.bn.weight, self.bn.bias]
        bufs = [dn_bn.running_mean, dn_bn.running_var, self.bn.running_mean, self.bn.running_var]
        return AF.gcn0(x, self.A, self.training, self.bn.momentum, self.bn.eps, self.PA, params, bufs)

    def forward_tokens(self, tok, dims):
        self._bump()
        params = self._abd() + [self.bn.weight, self.bn.bias]
        bufs = [self.bn.running_mean, self.bn.running_var]
        if self.has_down:
            dn_conv, dn_bn = self.down[0], self.down[1]
            params += [dn_conv.weight, dn_conv.bias, dn_bn.weight, dn_bn.bias]
            bufs += [dn_bn.running_mean, dn_bn.running_var]
        return AF.agcn(tok, dims, self.A, self.training, self.bn.momentum, self.bn.eps, self.PA, params, bufs, self.has_down)

    def forward(self, x):
        N, C, T, V = x.shape
        if not x.is_cuda:
            raise RuntimeError("altformer_b200 modules run on CUDA tensors only (there is no CPU fallback)")
        dims = (N, T, V)
        if C == 3 and self.has_down:
            skel = x.permute(0, 2, 3, 1)
            skel = skel if skel.is_contiguous() else skel.contiguous()
            if skel.dtype != torch.float32:
                skel = skel.float()
            return from_tokens(self.forward_skeleton(skel), dims)
        if C % 8 != 0:
            raise RuntimeError("altformer_b200.unit_agcn: in_channels must be 3 or a multiple of 8")
        tok, dims = to_tokens(x)
        return from_tokens(self.forward_tokens(tok, dims), dims)
