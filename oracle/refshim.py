"""Import the *unmodified* reference modules from /root/reference on CPU.  TEST INFRASTRUCTURE ONLY.

Two shims:
  1. `timm.models.layers` is not installed; the reference only uses `DropPath`
     (model/AltFormer/model_ST.py:10,79).  We inject a stub with timm-0.9.12 semantics.
  2. `unit_agcn.forward` calls `self.A.cuda(x.get_device())` (model/unit_agcn.py:75), which raises
     on CPU tensors.  While a reference module runs we patch `Tensor.cuda` to be the identity.
The reference directory does not exist on the GPU box, so everything here is guarded by
`available()`; nothing under `tests -m gpu`, `smoke()` or `bench.py` depends on it.
"""
import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("ALTFORMER_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "unit_agcn.py"))


def _install_timm_stub():
    if "timm.models.layers" in sys.modules:
        return

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = lambda v: (v, v)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


@contextlib.contextmanager
def cpu_cuda_noop():
    orig = torch.Tensor.cuda

    def fake(self, *a, **k):
        return self

    torch.Tensor.cuda = fake
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def load():
    """Returns a namespace with the reference classes."""
    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    _install_timm_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.unit_agcn = importlib.import_module("model.unit_agcn").unit_agcn
    ns.Unit2D = importlib.import_module("model.net").Unit2D
    st = importlib.import_module("model.AltFormer.model_ST")
    ts = importlib.import_module("model.AltFormer.model_TS")
    ns.ST, ns.TS, ns.Block, ns.Attention, ns.Mlp = st.ST, ts.TS, st.Block, st.Attention, st.Mlp
    ns.ST_GCN_AltFormer = importlib.import_module("model.AltFormer.ST_GCN_AltFormer").ST_GCN_AltFormer
    ns.graph = importlib.import_module("graph")
    return ns


def set_identity_droppath(module):
    for name, child in module.named_children():
        if type(child).__name__ == "DropPath":
            setattr(module, name, torch.nn.Identity())
        else:
            set_identity_droppath(child)
    return module


class _CpuReference(torch.nn.Module):
    """The unmodified reference ST_GCN_AltFormer running on CPU tensors (Tensor.cuda patched to identity per call)."""

    def __init__(self, mod):
        super().__init__()
        self.mod = mod

    def forward(self, x):
        with cpu_cuda_noop():
            return self.mod(x)


def build_model(state, num_class, T, V, style, graph, train=True):
    """Reference ST_GCN_AltFormer (model/AltFormer/ST_GCN_AltFormer.py:16-87) with `state` loaded, adjacency de-aliased
    (SURVEY 8c shim 2) and DropPath as constructed (drop_path_rate 0.1: the reference's own training behaviour).  Used by
    bench.py's CPU arm when the reference tree is present, so that arm times the reference's modules, not the port."""
    ref = load()
    mod = ref.ST_GCN_AltFormer(channel=3, num_class=num_class, num_frame=T, num_joints=V, style=style, graph=graph,
                               graph_args={"labeling_mode": "spatial"})
    A = torch.from_numpy((ref.graph.SHRE if V == 22 else ref.graph.LMDHG)("spatial").A).float()
    mod.gcn0.A = A.clone()
    missing, unexpected = mod.load_state_dict(state, strict=False)
    if unexpected or any(not k.startswith(("modelA.", "modelB.")) for k in missing):
        raise RuntimeError(f"reference state_dict mismatch: missing {missing[:4]}, unexpected {unexpected[:4]}")
    return _CpuReference(mod).train(train)
