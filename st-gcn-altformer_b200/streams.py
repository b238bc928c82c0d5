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


def random_augment(x, generator=None):
    """The reference's random policy drawn on the device: one of the four transforms per sample with equal probability
    (`randint(0, 3)`, :146-147), factor ~ U(0.8, 1.2) (:87-90), offsets ~ U(-0.1, 0.1) (:99-101, :110-118), four distinct
    joints (`shuffle(all_joint)[0:4]`, :113-115), r ~ U(0, 1) (:129).  Returns (augmented batch, kind, params)."""
    N, _, V, _ = x.shape
    dev = x.device
    u = torch.rand((N, 16), device=dev, generator=generator)
    kind = torch.randint(0, 4, (N,), device=dev, generator=generator, dtype=torch.int32)
    joints = torch.rand((N, V), device=dev, generator=generator).argsort(dim=1)[:, :4].float()
    params = torch.empty((N, 16), device=dev, dtype=torch.float32)
    k = kind[:, None]
    scale = torch.cat([0.8 + 0.4 * u[:, :1], torch.zeros((N, 15), device=dev)], 1)
    shift = torch.cat([-0.1 + 0.2 * u[:, :3], torch.zeros((N, 13), device=dev)], 1)
    noise = torch.cat([joints, -0.1 + 0.2 * u[:, 4:16]], 1)
    tint = torch.cat([u[:, :1], torch.zeros((N, 15), device=dev)], 1)
    params = torch.where(k == AUG_SCALE, scale, torch.where(k == AUG_SHIFT, shift, torch.where(k == AUG_NOISE, noise, tint)))
    return augment(x, kind, params.contiguous()), kind, params


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
