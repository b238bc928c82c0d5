"""Host-logic dry run (CPU): every kernel launch is replaced by an arity/type check against the
ctypes signatures of the C ABI, so the Python plumbing (autograd Functions, module wiring, argument
order of each afb_* call, flat-buffer trainer) is exercised without a GPU.  No numerics are checked
here -- outputs are uninitialised memory by construction."""
import ctypes

import pytest
import torch

import altformer_b200 as ab
from altformer_b200 import _lib, functional as AF, ops
from oracle import altformer_oracle as O


class FakeLib:
    def __init__(self, real):
        self.sig = real._afb_signatures
        self.calls = []

    def __getattr__(self, name):
        if name not in self.sig:
            raise AttributeError(name)
        want = self.sig[name]

        def fn(*args):
            assert len(args) == len(want), f"{name}: {len(args)} args, header declares {len(want)}"
            for i, (a, t) in enumerate(zip(args, want)):
                if t is ctypes.c_void_p:
                    assert a is None or isinstance(a, int), f"{name} arg {i}: pointer expected, got {type(a)}"
                elif t in (ctypes.c_int32, ctypes.c_int64):
                    assert isinstance(a, int) and not isinstance(a, bool) or isinstance(a, bool), f"{name} arg {i}: int expected, got {a!r}"
                elif t is ctypes.c_float:
                    assert isinstance(a, (int, float)), f"{name} arg {i}: float expected, got {a!r}"
                else:  # POINTER(struct): must be a byref of the right struct
                    assert "CArgObject" in type(a).__name__, f"{name} arg {i}: byref(struct) expected"
            self.calls.append(name)
            return 0
        return fn


@pytest.fixture()
def dry(monkeypatch):
    fake = FakeLib(_lib.lib())
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(ops, "stream", lambda: 0)
    monkeypatch.setattr(ops, "ensure_device", lambda t: None)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    yield fake
    AF.set_precision("bf16")


def _model(style, T=8, V=22, cls=14):
    torch.manual_seed(0)
    return ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style=style, graph="graph.SHRE" if V == 22 else "graph.LMDHG",
                               graph_args={"labeling_mode": "spatial"})


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
@pytest.mark.parametrize("style", ["ST", "TS", None])
def test_model_forward_backward_plumbing(dry, mode, style):
    AF.set_precision(mode)
    m = _model(style)
    x, y = O.synthetic_batch(2, 8, 22, 14)
    logits = m(x)
    assert logits.shape == (2, 14) and logits.dtype == torch.float32
    loss = AF.cross_entropy(logits, y)
    loss.backward()
    live = {n for n, _ in m.live_parameters()}
    got = {n for n, p in m.named_parameters() if p.grad is not None}
    assert got == live, (sorted(live - got)[:5], sorted(got - live)[:5])
    for name in ("afb_gcn0_fwd", "afb_gcn0_bwd", "afb_gemm_tn", "afb_gemm_dw", "afb_attention_fwd", "afb_attention_bwd",
                 "afb_layernorm_bwd", "afb_bn_bwd_apply", "afb_softmax_ce"):
        assert name in dry.calls, name


def test_eval_forward_lmdhg(dry):
    m = _model("ST", T=8, V=46).eval()
    x, _ = O.synthetic_batch(1, 8, 46, 14)
    with torch.no_grad():
        assert m(x).shape == (1, 14)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_standalone_modules(dry, mode):
    AF.set_precision(mode)
    A = O.spatial_graph(22)
    x = torch.randn(2, 64, 8, 22, requires_grad=True)
    for mod in (ab.unit_agcn(64, 64, A), ab.unit_agcn(64, 128, A), ab.TCN_GCN_unit(64, 64, A, dropout=0.0), ab.Unit2D(64, 64, 9)):
        y = mod(x)
        assert y.shape[0] == 2 and y.shape[2:] == (8, 22)
        y.float().sum().backward()
        assert x.grad is not None and x.grad.shape == x.shape
        assert all(p.grad is not None for n, p in mod.named_parameters()), [n for n, p in mod.named_parameters() if p.grad is None]
    t = torch.randn(3, 22, 256, requires_grad=True)
    blk = ab.Block(256, 8, mlp_ratio=2.0, qkv_bias=True, drop_path=0.1)
    for mod in (blk, blk.attn, blk.mlp):
        out = mod(t)
        assert out.shape == t.shape
        out.float().sum().backward()
    st = ab.ST(14, num_frame=8, num_joints=22, depth=1)
    assert st(torch.randn(2, 128, 8, 22)).shape == (2, 14)
    ts = ab.TS(14, num_frame=8, num_joints=22, depth=1)
    assert ts(torch.randn(2, 128, 8, 22)).shape == (2, 14)


def test_trainer_step_plumbing(dry):
    m = _model("ST")
    tr = ab.DataParallelTrainer(m, use_graph=False)
    assert tr.layout.total % 8 == 0 and tr.layout.total >= sum(p.numel() for _, p in m.live_parameters())
    for p in tr.params:
        assert p.data_ptr() % 16 == 0 and p._afb_shadow.data_ptr() % 16 == 0 and p.grad is p._afb_grad
    x, y = O.synthetic_batch(2, 8, 22, 14)
    loss, logits = tr.step(x, y)
    assert logits.shape == (2, 14)
    assert "afb_adamw" in dry.calls and "afb_step_inc" in dry.calls


def test_errors_without_gpu():
    """No CPU fallback: the real library refuses CPU tensors / the modules refuse to run."""
    m = _model("ST")
    x, _ = O.synthetic_batch(2, 8, 22, 14)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cast(torch.zeros(4), torch.bfloat16)
    with pytest.raises(RuntimeError):
        ab.Unit2D(64, 64, 9)(torch.zeros(1, 64, 4, 22))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours): exactly one JSON line on stdout."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
