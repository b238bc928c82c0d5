"""Input streams and ensemble combine on device.

palm   = x - x[frame 0, joint 1]   data_process/Hand_Dataset.py:61
bone   = joint - parent joint      data_process/Hand_Dataset.py:200-217 (table :201-202)
motion = next frame - this frame   data_process/Hand_Dataset.py:183-198 (last frame zero)
combine 0.8*ST + 0.2*TS            SHREC/ST_TS/emsemble.py:217-218
augment = scale | shift | noise | time_interpolate, one per sample   data_process/Hand_Dataset.py:84-157
"""
import torch

from . import ops
from .graph.skeletons import SHREC_PARENT

_parent_cache = {}


def bone(x, parent=SHREC_PARENT):
    """x (N, T, V, 3) float32 CUDA."""
    key = (x.device, tuple(parent))
    if key not in _parent_cache:
        _parent_cache[key] = torch.tensor(parent, dtype=torch.int32, device=x.device)
    return ops.bone_stream(x.contiguous(), _parent_cache[key])


def palm_normalise(x, joint=1):
    """x (N, T, V, 3) float32 CUDA -> x - x[:, :1, joint:joint+1, :]  (Hand_Dataset.py:61: `skeleton -= skeleton[0][1]`)."""
    return ops.palm_center(x.contiguous(), joint)


def motion(x):
    return ops.motion_stream(x.contiguous())


AUG_SCALE, AUG_SHIFT, AUG_NOISE, AUG_TIME = 0, 1, 2, 3


def augment(x, kind, params):
    """Hand_Dataset.data_aug (data_process/Hand_Dataset.py:84-157) for a whole batch in one launch, parameters explicit.
    x (N, T, V, 3) fp32 CUDA; kind (N,) int32 in {AUG_SCALE, AUG_SHIFT, AUG_NOISE, AUG_TIME} (anything else: copy);
    params (N, 16) fp32: [0] factor (scale) | [0:3] offset (shift) | [0:4] joint ids + [4:16] their xyz offsets (noise) |
    [0] r (time_interpolate)."""
    return ops.augment(x.contiguous(), kind, params)


def draw_augment_params(N, V, device, generator=None):
    """The reference's random policy as tensors (pure torch, any device): one of the four transforms per sample with equal
    probability (`randint(0, 3)`, Hand_Dataset.py:146-147), factor ~ U(0.8, 1.2) (:87-90), offsets ~ U(-0.1, 0.1) (:99-101,
    :110-118), four distinct joints (`shuffle(all_joint)[0:4]`, :113-115), r ~ U(0, 1) (:129).
    Returns (kind (N,) int32, params (N, 16) float32) in the layout afb_augment takes."""
    u = torch.rand((N, 16), device=device, generator=generator)
    kind = torch.randint(0, 4, (N,), device=device, generator=generator, dtype=torch.int32)
    joints = torch.rand((N, V), device=device, generator=generator).argsort(dim=1)[:, :4].float()
    zeros = lambda c: torch.zeros((N, c), device=device)  # noqa: E731
    k = kind[:, None]
    scale = torch.cat([0.8 + 0.4 * u[:, :1], zeros(15)], 1)
    shift = torch.cat([-0.1 + 0.2 * u[:, :3], zeros(13)], 1)
    noise = torch.cat([joints, -0.1 + 0.2 * u[:, 4:16]], 1)
    tint = torch.cat([u[:, :1], zeros(15)], 1)
    params = torch.where(k == AUG_SCALE, scale, torch.where(k == AUG_SHIFT, shift, torch.where(k == AUG_NOISE, noise, tint)))
    return kind, params.contiguous()


def random_augment(x, generator=None):
    """Hand_Dataset.data_aug with its random policy drawn on the device (draw_augment_params) and applied in one launch.
    Returns (augmented batch, kind, params)."""
    N, _, V, _ = x.shape
    kind, params = draw_augment_params(N, V, x.device, generator)
    return augment(x, kind, params), kind, params


def combine(logits_st, logits_ts, w_st=0.8, w_ts=0.2):
    return ops.axpby(logits_st.contiguous(), w_st, logits_ts.contiguous(), w_ts)


@torch.no_grad()
def ensemble_forward(x, models):
    """models: dict stream -> (model_ST, model_TS) with stream in {'joint','bone','motion'}.
    Returns the sum over streams of 0.8*ST + 0.2*TS logits (emsemble.py:215-225 semantics)."""
    total = None
    for name, (m_st, m_ts) in models.items():
        xs = x if name == "joint" else (bone(x) if name == "bone" else motion(x))
        s = combine(m_st(xs), m_ts(xs))
        total = s if total is None else ops.axpby(total, 1.0, s, 1.0)
    return total
