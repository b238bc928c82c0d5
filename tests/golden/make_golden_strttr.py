"""Golden vectors for the STR / TTR ablation models (round 2; run here, where /root/reference exists).

  strttr_STR_22.pt / strttr_TTR_22.pt   the UNMODIFIED reference STR_TTR (STR_TTR/STR_TTR.py, STR.py, TTR.py) in train mode
      (batch-statistics BatchNorm, DropPath replaced by identity as in the other goldens) with a seeded state_dict:
      forward output and the gradients of sum(out * w) w.r.t. every parameter that receives one.
Shims (import-time only, the model code is untouched): timm.data / timm.models.helpers / timm.models.registry names the files
import but never call, `visualizer.get_local` (a pass-through decorator), and the adjacency de-aliasing of the other goldens.

    python tests/golden/make_golden_strttr.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import altformer_oracle as O  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def compact(case, full_limit=8192, nsample=512):
    """Keep small tensors whole; for large ones keep the norm and a strided sample (same format as make_golden.py,
    read by tests/goldenlib.check_entry)."""
    out = {}
    for k, v in case.items():
        if v.numel() <= full_limit:
            out[k] = v
        else:
            flat = v.reshape(-1)
            stride = max(flat.numel() // nsample, 1)
            out[k] = {"shape": tuple(v.shape), "norm": float(flat.double().norm()), "stride": stride, "sample": flat[::stride].clone()}
    return out


def _stubs():
    ref = refshim.load()
    data = types.ModuleType("timm.data")
    data.IMAGENET_DEFAULT_MEAN = data.IMAGENET_DEFAULT_STD = (0.0, 0.0, 0.0)
    helpers = types.ModuleType("timm.models.helpers")
    helpers.load_pretrained = lambda *a, **k: None
    registry = types.ModuleType("timm.models.registry")
    registry.register_model = lambda f: f
    vis = types.ModuleType("visualizer")
    vis.get_local = lambda *a, **k: (lambda f: f)
    sys.modules.update({"timm.data": data, "timm.models.helpers": helpers, "timm.models.registry": registry, "visualizer": vis})
    sys.path.insert(0, os.path.join(refshim.REF_ROOT, "STR_TTR"))
    import importlib
    return ref, importlib.import_module("STR_TTR").STR_TTR


def main():
    ref, STR_TTR = _stubs()
    N, T, V, cls = 3, 8, 22, 14
    A = O.spatial_graph(V)
    for style, seed in (("STR", 81), ("TTR", 82)):
        mod = STR_TTR(channel=3, num_class=cls, num_frame=T, num_joints=V, style=style, graph="graph.SHRE",
                      graph_args={"labeling_mode": "spatial"})
        mod.gcn.A = A.clone()                       # de-alias the adjacency (SURVEY 8c shim 2, as in make_golden.py)
        spec = O.strttr_spec(style, 3, cls, T, V)
        assert set(spec) == set(mod.state_dict()) and all(tuple(mod.state_dict()[k].shape) == tuple(v) for k, v in spec.items())
        st = O.random_state(spec, seed)
        mod.load_state_dict(st, strict=True)
        refshim.set_identity_droppath(mod)
        mod.train()
        x, _ = O.synthetic_batch(N, T, V, cls, seed + 100)
        with refshim.cpu_cuda_noop():
            out = mod(x)
        g = torch.Generator().manual_seed(seed + 7)
        w = torch.randn(out.shape, generator=g)
        (out * w).sum().backward()
        grads = {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
        torch.save({"y": out.detach().clone(), "w": w, "grads": compact(grads), "state_seed": seed, "batch_seed": seed + 100,
                    "shape": (N, T, V, cls), "style": style}, os.path.join(OUT, f"strttr_{style}_22.pt"))
        print(f"strttr_{style}_22: out", tuple(out.shape), float(out.norm()), "grads", len(grads))


if __name__ == "__main__":
    main()
