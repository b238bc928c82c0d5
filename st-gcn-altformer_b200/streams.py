"""Input streams and ensemble combine on device.

palm   = x - x[frame 0, joint 1]   data_process/Hand_Dataset.py:61
bone   = joint - parent joint      data_process/Hand_Dataset.py:200-217 (table :201-202)
motion = next frame - this frame   data_process/Hand_Dataset.py:183-198 (last frame zero)
combine 0.8*ST + 0.2*TS            SHREC/ST_TS/emsemble.py:217-218
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
