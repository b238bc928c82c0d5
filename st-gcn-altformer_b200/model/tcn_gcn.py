"""TCN_GCN_unit (unit_agcn branch): tcn1(gcn1(x)) + x, the layer of the "unit_agcn + temporal-conv
stack" sweep.  Reference: model/ST_TR/ST_TR_new.py:355-385 incl. the strided / channel-changing variant with its
`down1 = Unit2D(k=1, stride)` skip path (:369-374); the gcn_unit_attention branch is outside the built scope.  The
residual add is fused into the BN+ReLU pass of tcn1."""
import torch.nn as nn

from .net import Unit2D
from .unit_agcn import unit_agcn
from ._tokens import from_tokens, to_tokens


class TCN_GCN_unit(nn.Module):
    # keyword arguments of the reference constructor (ST_TR_new.py:300-340) that only configure its gcn_unit_attention branch or
    # bookkeeping; accepted (with the reference's defaults) so call sites port unchanged, anything else is an error
    _REFERENCE_ONLY = ("attention", "only_attention", "tcn_attention", "only_temporal_attention", "attention_3", "relative", "weight_matrix",
                       "device", "more_channels", "drop_connect", "data_normalization", "skip_conn", "adjacency", "starting_ch",
                       "visualization", "all_layers", "dv", "dk", "Nh", "num", "dim_block1", "dim_block2", "dim_block3", "num_point",
                       "layer", "last", "last_graph", "agcn", "bn_flag")

    def __init__(self, in_channel, out_channel, A, kernel_size=9, stride=1, dropout=0.5, use_local_bn=False,
                 mask_learning=False, **reference_only):
        super().__init__()
        unknown = [k for k in reference_only if k not in self._REFERENCE_ONLY]
        if unknown:
            raise TypeError(f"altformer_b200.TCN_GCN_unit: unexpected keyword argument(s) {unknown}")
        if reference_only.get("attention") or reference_only.get("tcn_attention"):
            raise ValueError("altformer_b200.TCN_GCN_unit: the gcn_unit_attention / temporal-attention branches are outside the built scope")
        self.gcn1 = unit_agcn(in_channel, out_channel, A, use_local_bn=use_local_bn, mask_learning=mask_learning)
        self.tcn1 = Unit2D(out_channel, out_channel, kernel_size=kernel_size, dropout=dropout, stride=stride)
        # ST_TR_new.py:369-374: a 1 x 1 (strided) Unit2D on the skip path when the shape changes
        self.down1 = Unit2D(in_channel, out_channel, kernel_size=1, stride=stride) if (in_channel != out_channel or stride != 1) else None

    def out_dims(self, dims):
        return self.tcn1.out_dims(dims)

    def forward_tokens(self, tok, dims):
        res = tok if self.down1 is None else self.down1.forward_tokens(tok, dims)
        return self.tcn1.forward_tokens(self.gcn1.forward_tokens(tok, dims), dims, res_post=res)

    def forward(self, x):
        tok, dims = to_tokens(x)
        return from_tokens(self.forward_tokens(tok, dims), self.out_dims(dims))
