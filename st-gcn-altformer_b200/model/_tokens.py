"""Layout helpers.  Inside the package every activation is a channels-last token matrix
[N*T*V, C]; at module boundaries tensors keep the reference's logical (N, C, T, V) shape but are
physically channels-last, so `(b c f p) -> ((b f) p) c` is a view, never a copy."""
import torch

from altformer_b200 import functional as AF


def to_tokens(x):
    """logical (N, C, T, V) -> ([N*T*V, C] contiguous, (N, T, V)).  Free when x is channels-last."""
    if x.dim() != 4:
        raise RuntimeError(f"expected a (N, C, T, V) tensor, got shape {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("altformer_b200 modules run on CUDA tensors only (there is no CPU fallback)")
    N, C, T, V = x.shape
    nhwc = x.permute(0, 2, 3, 1)
    if not nhwc.is_contiguous():
        nhwc = nhwc.contiguous()  # boundary layout conversion for NCHW-contiguous callers
    return AF.to_act(nhwc.view(N * T * V, C)), (N, T, V)


def from_tokens(y, dims):
    """[N*T*V, C] -> logical (N, C, T, V) view (channels-last memory)."""
    N, T, V = dims
    return y.view(N, T, V, y.shape[-1]).permute(0, 3, 1, 2)
