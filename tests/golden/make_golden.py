"""Generate golden vectors by running the UNMODIFIED reference modules (CPU, fp32).

Run here (the authoring container, where /root/reference exists):
    python tests/golden/make_golden.py
Weights/inputs are regenerated from seeds by `oracle.altformer_oracle.random_state`, so only the
reference's *outputs and gradients* are stored.  The GPU box never runs this script.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import altformer_oracle as O  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(0)
torch.set_num_threads(8)


def cot_like(t, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(t.shape, generator=g)


def grads_of(module, names):
    sd = dict(module.named_parameters())
    return {k: (sd[k].grad.clone() if sd[k].grad is not None else None) for k in names if k in sd}


def run_module(mod, state, x, training, need_dx=False, seed=7):
    missing = mod.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    mod.train(training)
    x = x.clone().requires_grad_(need_dx)
    with refshim.cpu_cuda_noop():
        y = mod(x)
        cot = cot_like(y, seed)
        (y * cot).sum().backward()
    out = {"y": y.detach().clone()}
    if need_dx:
        out["dx"] = x.grad.clone()
    for k, p in mod.named_parameters():
        if p.grad is not None:
            out["grad." + k] = p.grad.clone()
    for k, b in mod.named_buffers():
        if "running" in k:
            out["buf." + k] = b.clone()
    return out


def compact(case, full_limit=8192, nsample=512):
    """Keep small tensors whole; for large ones keep the norm and a strided sample."""
    out = {}
    for k, v in case.items():
        if not torch.is_tensor(v) or v.numel() <= full_limit:
            out[k] = v
        else:
            flat = v.reshape(-1)
            stride = max(flat.numel() // nsample, 1)
            out[k] = {"shape": tuple(v.shape), "norm": float(flat.double().norm()), "stride": stride,
                      "sample": flat[::stride].clone()}
    return out


def main():
    ref = refshim.load()
    meta = {}

    # ---- graphs ------------------------------------------------------------------------
    g = {"SHRE": torch.from_numpy(ref.graph.SHRE("spatial").A).float(),
         "LMDHG": torch.from_numpy(ref.graph.LMDHG("spatial").A).float()}
    torch.save(g, os.path.join(OUT, "graphs.pt"))

    A22, A46 = g["SHRE"], g["LMDHG"]

    def agcn_case(cin, cout, N, T, V, A, training, seed, need_dx):
        st = O.random_state(O.agcn_spec("", cin, cout, V), seed)
        mod = ref.unit_agcn(cin, cout, A.clone())
        mod.A = A.clone()  # de-alias (SURVEY §8c shim 2)
        gx = torch.Generator().manual_seed(seed + 100)
        x = 0.5 * torch.randn(N, cin, T, V, generator=gx)
        return run_module(mod, st, x, training, need_dx)

    cases = {}
    cases["agcn_3_128_train"] = agcn_case(3, 128, 2, 8, 22, A22, True, 11, False)
    cases["agcn_3_128_eval"] = agcn_case(3, 128, 2, 8, 22, A22, False, 11, False)
    cases["agcn_3_128_v46_train"] = agcn_case(3, 128, 2, 6, 46, A46, True, 12, False)
    cases["agcn_64_64_train"] = agcn_case(64, 64, 2, 8, 22, A22, True, 13, True)
    cases["agcn_32_64_train"] = agcn_case(32, 64, 2, 8, 22, A22, True, 14, True)

    def unit2d_case(cin, cout, k, N, T, V, training, seed):
        st = O.random_state(O.unit2d_spec("", cin, cout, k), seed)
        mod = ref.Unit2D(cin, cout, kernel_size=k)
        gx = torch.Generator().manual_seed(seed + 100)
        x = torch.randn(N, cin, T, V, generator=gx)
        return run_module(mod, st, x, training, True)

    cases["unit2d_32_train"] = unit2d_case(32, 32, 9, 2, 12, 22, True, 21)
    cases["unit2d_32_eval"] = unit2d_case(32, 32, 9, 2, 12, 22, False, 21)

    # TCN_GCN_unit (agcn branch) restated from its two members + residual (ST_TR_new.py:376-385);
    # the class itself needs 30 unrelated ctor args, so compose the reference members directly.
    def tcn_gcn_case(C, N, T, V, A, seed):
        spec = O.OrderedDict()
        spec.update(O.agcn_spec("gcn1.", C, C, V))
        spec.update(O.unit2d_spec("tcn1.", C, C, 9))
        st = O.random_state(spec, seed)

        class Unit(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.gcn1 = ref.unit_agcn(C, C, A.clone())
                self.gcn1.A = A.clone()
                self.tcn1 = ref.Unit2D(C, C, kernel_size=9, dropout=0.0)

            def forward(self, x):
                return self.tcn1(self.gcn1(x)) + x

        gx = torch.Generator().manual_seed(seed + 100)
        x = torch.randn(N, C, T, V, generator=gx)
        return run_module(Unit(), st, x, True, True)

    cases["tcn_gcn_64_train"] = tcn_gcn_case(64, 2, 8, 22, A22, 31)

    def block_case(D, B, L, seed):
        st = O.random_state(O.block_spec("", D), seed)
        mod = ref.Block(dim=D, num_heads=8, mlp_ratio=2.0, qkv_bias=True, drop_path=0.0,
                        norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6))
        gx = torch.Generator().manual_seed(seed + 100)
        x = torch.randn(B, L, D, generator=gx)
        return run_module(mod, st, x, True, True)

    cases["block_64_L22"] = block_case(64, 3, 22, 41)
    cases["block_128_L8"] = block_case(128, 2, 8, 42)

    def stage_case(kind, N, T, V, cls, seed, cin=32, d1=64, depth=2):
        spec = (O.st_spec if kind == "ST" else O.ts_spec)("", cls, T, V, cin, d1, depth)
        # fcn is hard-wired to 512 input channels in the reference
        st = O.random_state(spec, seed)
        cls_ = ref.ST if kind == "ST" else ref.TS
        mod = cls_(cls, num_frame=T, num_joints=V, in_chans=cin, embed_dim_ratio=d1, depth=depth,
                   num_heads=8, mlp_ratio=2.0, qkv_bias=True, drop_path_rate=0.0)
        gx = torch.Generator().manual_seed(seed + 100)
        x = torch.randn(N, cin, T, V, generator=gx)
        return run_module(mod, st, x, True, True)

    cases["st_small"] = stage_case("ST", 2, 8, 22, 14, 51)
    cases["ts_small"] = stage_case("TS", 2, 8, 22, 14, 52)

    def model_case(style, N, T, V, cls, graph, seed, training=True):
        st = O.random_state(O.model_spec(3, cls, T, V), seed)
        mod = ref.ST_GCN_AltFormer(channel=3, num_class=cls, num_frame=T, num_joints=V, style=style,
                                   graph=graph, graph_args={"labeling_mode": "spatial"})
        A = A22 if V == 22 else A46
        mod.gcn0.A = A.clone()
        refshim.set_identity_droppath(mod)
        x, _ = O.synthetic_batch(N, T, V, cls, seed + 100)
        full = run_module(mod, st, x, training, False)
        # the full gradient set is 64 MB: keep logits, all gradient norms and the small tensors
        keep = {"y": full["y"]}
        norms = {}
        for k, v in full.items():
            if k.startswith("grad."):
                norms[k[5:]] = float(v.double().norm())
                if v.numel() <= 4096:
                    keep[k] = v
            elif k.startswith("buf.") and ("gcn0" in k or "tcn0" in k):
                keep[k] = v
        keep["grad_norms"] = norms
        return keep

    cases["model_ST_22"] = model_case("ST", 2, 8, 22, 14, "graph.SHRE", 61)
    cases["model_TS_22"] = model_case("TS", 2, 8, 22, 14, "graph.SHRE", 62)
    cases["model_both_22"] = model_case(None, 2, 8, 22, 28, "graph.SHRE", 63)
    cases["model_ST_46_eval"] = model_case("ST", 1, 8, 46, 14, "graph.LMDHG", 64, training=False)

    # ---- state_dict key/shape contract --------------------------------------------------
    mod = ref.ST_GCN_AltFormer(channel=3, num_class=28, num_frame=32, num_joints=22, style="ST",
                               graph="graph.SHRE", graph_args={"labeling_mode": "spatial"})
    meta["state_keys"] = [(k, tuple(v.shape)) for k, v in mod.state_dict().items()]
    meta["n_params"] = sum(p.numel() for p in mod.parameters())

    total = 0
    for name, c in cases.items():
        c = compact(c)
        path = os.path.join(OUT, name + ".pt")
        torch.save(c, path)
        total += os.path.getsize(path)
        print(f"{name:28s} {os.path.getsize(path) / 1024:8.1f} KB")
    torch.save(meta, os.path.join(OUT, "meta.pt"))
    print("total KB", total / 1024)


if __name__ == "__main__":
    main()
