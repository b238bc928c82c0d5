"""Pin the CPU oracle against golden vectors produced by the unmodified reference modules."""
import pytest
import torch

from oracle import altformer_oracle as O
from tests import goldenlib as G

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order


def _cot(t, seed=7):
    return torch.randn(t.shape, generator=torch.Generator().manual_seed(seed))


def _run(fn, state, x, need_dx):
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in state.items()}
    x = x.clone().requires_grad_(need_dx)
    y = fn(x, params)
    (y * _cot(y)).sum().backward()
    return y, x.grad, params


def _compare(case, y, dx, params, tol=TOL):
    G.check_entry(case["y"], y, tol, "y")
    if "dx" in case:
        G.check_entry(case["dx"], dx, tol, "dx")
    n = 0
    for k, e in case.items():
        if k.startswith("grad."):
            g = params[k[5:]].grad
            ref_norm = e["norm"] if isinstance(e, dict) else float(e.double().norm())
            if ref_norm < 1e-4:  # analytically-zero gradients: absolute check
                assert g is None or float(g.norm()) < 1e-3, k
            else:
                G.check_entry(e, g, 20 * tol, k)
            n += 1
        elif k.startswith("buf."):
            G.check_entry(e, params[k[4:]], tol, k)
    return n


def test_graphs():
    g = G.load("graphs")
    assert torch.equal(O.spatial_graph("SHRE"), g["SHRE"])
    assert torch.equal(O.spatial_graph("LMDHG"), g["LMDHG"])


@pytest.mark.parametrize("name,cin,cout,N,T,V,train,seed,dx", [
    ("agcn_3_128_train", 3, 128, 2, 8, 22, True, 11, False),
    ("agcn_3_128_eval", 3, 128, 2, 8, 22, False, 11, False),
    ("agcn_3_128_v46_train", 3, 128, 2, 6, 46, True, 12, False),
    ("agcn_64_64_train", 64, 64, 2, 8, 22, True, 13, True),
    ("agcn_32_64_train", 32, 64, 2, 8, 22, True, 14, True),
])
def test_agcn(name, cin, cout, N, T, V, train, seed, dx):
    A = O.spatial_graph(V)
    st = O.random_state(O.agcn_spec("", cin, cout, V), seed)
    x = 0.5 * torch.randn(N, cin, T, V, generator=torch.Generator().manual_seed(seed + 100))
    y, gx, params = _run(lambda x, p: O.agcn_forward(x, p, "", A, train), st, x, dx)
    assert _compare(G.load(name), y, gx, params) >= 18


@pytest.mark.parametrize("name,train", [("unit2d_32_train", True), ("unit2d_32_eval", False)])
def test_unit2d(name, train):
    st = O.random_state(O.unit2d_spec("", 32, 32, 9), 21)
    x = torch.randn(2, 32, 12, 22, generator=torch.Generator().manual_seed(121))
    y, gx, params = _run(lambda x, p: O.unit2d_forward(x, p, "", train), st, x, True)
    _compare(G.load(name), y, gx, params)


def test_tcn_gcn_unit():
    spec = O.OrderedDict()
    spec.update(O.agcn_spec("gcn1.", 64, 64, 22))
    spec.update(O.unit2d_spec("tcn1.", 64, 64, 9))
    st = O.random_state(spec, 31)
    x = torch.randn(2, 64, 8, 22, generator=torch.Generator().manual_seed(131))
    A = O.spatial_graph(22)
    y, gx, params = _run(lambda x, p: O.tcn_gcn_forward(x, p, "", A, True), st, x, True)
    _compare(G.load("tcn_gcn_64_train"), y, gx, params)


@pytest.mark.parametrize("name,D,B,L,seed", [("block_64_L22", 64, 3, 22, 41), ("block_128_L8", 128, 2, 8, 42)])
def test_block(name, D, B, L, seed):
    st = O.random_state(O.block_spec("", D), seed)
    x = torch.randn(B, L, D, generator=torch.Generator().manual_seed(seed + 100))
    y, gx, params = _run(lambda x, p: O.block_forward(x, p, ""), st, x, True)
    _compare(G.load(name), y, gx, params)


@pytest.mark.parametrize("kind,seed", [("ST", 51), ("TS", 52)])
def test_stage(kind, seed):
    spec = (O.st_spec if kind == "ST" else O.ts_spec)("", 14, 8, 22, 32, 64, 2)
    st = O.random_state(spec, seed)
    x = torch.randn(2, 32, 8, 22, generator=torch.Generator().manual_seed(seed + 100))
    fn = O.st_forward if kind == "ST" else O.ts_forward
    y, gx, params = _run(lambda x, p: fn(x, p, ""), st, x, True)
    _compare(G.load("st_small" if kind == "ST" else "ts_small"), y, gx, params)


@pytest.mark.parametrize("name,style,N,T,V,cls,seed,train", [
    ("model_ST_22", "ST", 2, 8, 22, 14, 61, True),
    ("model_TS_22", "TS", 2, 8, 22, 14, 62, True),
    ("model_both_22", None, 2, 8, 22, 28, 63, True),
    ("model_ST_46_eval", "ST", 1, 8, 46, 14, 64, False),
])
def test_model(name, style, N, T, V, cls, seed, train):
    case = G.load(name)
    st = O.random_state(O.model_spec(3, cls, T, V), seed)
    x, _ = O.synthetic_batch(N, T, V, cls, seed + 100)
    A = O.spatial_graph(V)
    y, _, params = _run(lambda x, p: O.model_forward(x, p, A, style, train), st, x, False)
    _compare(case, y, None, params, tol=5e-5)
    # every gradient norm the reference produced, and no gradient where it produced none
    for k, n in case["grad_norms"].items():
        g = params[k].grad
        assert g is not None, k
        if n > 1e-4:
            assert abs(float(g.double().norm()) - n) / n < 2e-3, (k, float(g.norm()), n)
    live = {k for k, v in params.items() if v.is_floating_point() and v.requires_grad and v.grad is not None}
    assert live == set(case["grad_norms"].keys())


def test_state_dict_contract():
    meta = G.load("meta")
    spec = O.model_spec(3, 28, 32, 22)
    assert [(k, tuple(s)) for k, s in spec.items()] == [(k, tuple(s)) for k, s in meta["state_keys"]]
    assert len(spec) == 366


def test_streams_match_reference_loops():
    x = torch.randn(2, 5, 22, 3)
    b = O.bone_stream(x)
    assert torch.allclose(b[0, 3, 7], x[0, 3, 7] - x[0, 3, 6])
    assert torch.allclose(b[:, :, 0], torch.zeros_like(b[:, :, 0]))
    m = O.motion_stream(x)
    assert torch.allclose(m[1, 2], x[1, 3] - x[1, 2]) and float(m[:, -1].abs().max()) == 0.0


def test_reference_checkpoint_semantics():
    """model_ST_22_refckpt: logits of the reference exactly as its constructor + .cuda() leave it (adjacency == 1e-6, see
    tests/golden/make_golden_extra.py).  The oracle reproduces them with A = 1e-6 -- and NOT with the graph adjacency."""
    case = G.load("model_ST_22_refckpt")
    N, T, V, cls = case["shape"]
    st = O.random_state(O.model_spec(3, cls, T, V), case["state_seed"])
    x, _ = O.synthetic_batch(N, T, V, cls, case["batch_seed"])
    y = O.model_forward(x, st, torch.full((3, V, V), 1e-6), "ST", False)
    G.check_entry(case["y"], y, 5e-5, "logits with the reference's effective adjacency")
    y_graph = O.model_forward(x, st, O.spatial_graph(V), "ST", False)
    assert float((y_graph - case["y"]).norm() / case["y"].norm()) > 1e-2


def test_streams_against_reference_hand_dataset():
    """bone / motion / palm normalisation vs the outputs of the reference's own Hand_Dataset methods (streams_22.pt)."""
    case = G.load("streams_22")
    x = case["x"][None]                      # (1, T, 22, 3)
    palm = x - x[:, :1, 1:2, :]
    assert torch.allclose(palm[0], case["palm"], atol=1e-6)
    assert torch.allclose(O.motion_stream(palm)[0], case["motion"], atol=1e-6)
    assert torch.allclose(O.bone_stream(palm)[0], case["bone"], atol=1e-6)


def test_augment_against_reference_data_aug():
    """The oracle's explicit-parameter augmentation vs Hand_Dataset.data_aug run on the reference class with its random draws
    recorded (augment_22.pt, tests/golden/make_golden_aug.py): scale, shift, noise on four joints, time_interpolate."""
    case = G.load("augment_22")
    y = O.augment(case["x"], case["kind"], case["params"])
    assert set(case["kind"].tolist()) == {0, 1, 2, 3}
    assert torch.allclose(y, case["y"], atol=1e-6)
    # an unknown kind leaves the sample untouched
    assert torch.equal(O.augment(case["x"][:1], torch.tensor([7]), case["params"][:1]), case["x"][:1])


@pytest.mark.parametrize("style", ["STR", "TTR"])
def test_strttr_against_reference(style):
    """The oracle's STR / TTR restatement (STR_TTR/STR.py:150-191, TTR.py:151-221, STR_TTR.py:62-84) vs the unmodified
    reference run in train mode (tests/golden/make_golden_strttr.py): output and every parameter gradient."""
    d = G.load(f"strttr_{style}_22")
    N, T, V, cls = d["shape"]
    st = O.random_state(O.strttr_spec(style, 3, cls, T, V), d["state_seed"])
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in st.items()}
    x, _ = O.synthetic_batch(N, T, V, cls, d["batch_seed"])
    y = O.strttr_forward(x, p, O.spatial_graph(V), style, True)
    G.check_entry(d["y"], y, 1e-5, f"{style} output")
    (y * d["w"]).sum().backward()
    assert len(d["grads"]) > 100
    for k, g in d["grads"].items():
        ref_norm = float(g.double().norm()) if torch.is_tensor(g) else g["norm"]
        if ref_norm > 1e-6:
            G.check_entry(g, p[k].grad, 1e-4, f"{style} grad {k}")


@pytest.mark.parametrize("name", ["unit2d_dim3_train", "unit2d_dim3_s2_train"])
def test_unit2d_dim3(name):
    """Unit2D(dim=3): 1 x k convolution along the joints (model/net.py:29-36), stride 1 and 2, vs the reference class."""
    case = G.load(name)
    cin, cout, k, N, T, V, stride = case["shape"]
    spec = O.unit2d_spec("", cin, cout, k)
    spec["conv.weight"] = (cout, cin, 1, k)
    st = O.random_state(spec, case["seed"])
    x = torch.randn(N, cin, T, V, generator=torch.Generator().manual_seed(case["seed"] + 100))
    y, gx, params = _run(lambda x, p: O.unit2d_forward(x, p, "", True, stride), st, x, True)
    _compare(case, y, gx, params)
