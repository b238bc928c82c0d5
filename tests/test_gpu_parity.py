"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the committed golden vectors
of the unmodified reference.  Tolerances: fp32 mode <= 1e-4, bf16 mode <= 1e-2 per kernel
on identical inputs, both as max-abs/max|ref| and relative L2; gradients of composite modules get a
small multiple; analytically-zero gradients are checked absolutely."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _diag():
    from tools import gpu_diag
    return gpu_diag


@pytest.mark.parametrize("group", ["simt", "elementwise", "gemm_tn", "gemm_mn", "gemm_dw", "attention", "gcn0", "modules", "model", "trainer"])
def test_kernels_against_reference_math(group):
    d = _diag()
    d.RESULTS.clear()
    d.run_group(group)
    failed = [n for n, ok in d.RESULTS if not ok]
    assert d.RESULTS and not failed, failed


def test_extension_is_loaded_not_a_fallback():
    import altformer_b200 as ab
    ab._lib.lib()
    maps = open("/proc/self/maps").read()
    assert "libaltformer_b200.so" in maps


@pytest.mark.parametrize("name,train", [("agcn_3_128_train", True), ("agcn_3_128_eval", False)])
def test_gcn0_against_reference_golden(name, train):
    """CUDA gcn0 (fp32 mode) vs the reference module's own output on the golden inputs."""
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    from tests import goldenlib as G
    case = G.load(name)
    ab.set_precision("fp32")
    try:
        A = O.spatial_graph(22)
        st = O.random_state(O.agcn_spec("", 3, 128, 22), 11)
        x = 0.5 * torch.randn(2, 3, 8, 22, generator=torch.Generator().manual_seed(111))
        mod = ab.unit_agcn(3, 128, A).cuda()
        mod.load_state_dict(st)
        mod.train(train)
        y = mod(x.cuda())
        G.check_entry(case["y"], y.float(), 1e-4, "y")
        if train:
            cot = torch.randn(y.shape, generator=torch.Generator().manual_seed(7)).cuda()
            (y.float() * cot).sum().backward()
            from tools.gpu_diag_modules import is_zero_class
            named = dict(mod.named_parameters())
            for k, p in named.items():
                ref = case["grad." + k]
                if is_zero_class(k, True):
                    # analytically zero (the reference only holds fp32 round-off there): absolute check against
                    # the scale of the sibling weight gradient
                    scale = float(named[k[:-4] + "weight"].grad.norm())
                    assert float(p.grad.norm()) <= 1e-3 * scale, (k, float(p.grad.norm()), scale)
                else:
                    G.check_entry(ref, p.grad, 5e-4, k)
    finally:
        ab.set_precision("bf16")


@pytest.mark.parametrize("name,style,cls,seed", [("model_ST_22", "ST", 14, 61), ("model_TS_22", "TS", 14, 62), ("model_both_22", None, 28, 63)])
def test_full_model_against_reference_golden(name, style, cls, seed):
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    from tests import goldenlib as G
    case = G.load(name)
    ab.set_precision("fp32")
    try:
        st = O.random_state(O.model_spec(3, cls, 8, 22), seed)
        x, _ = O.synthetic_batch(2, 8, 22, cls, seed + 100)
        mod = ab.ST_GCN_AltFormer(3, cls, num_frame=8, num_joints=22, style=style, graph="graph.SHRE", graph_args={"labeling_mode": "spatial"})
        mod.load_state_dict(st)
        mod = mod.cuda().train()
        for m in mod.modules():
            if type(m).__name__ == "DropPath":
                m.drop_prob = 0.0
        y = mod(x.cuda())
        G.check_entry(case["y"], y, 3e-4, "logits")
        cot = torch.randn(y.shape, generator=torch.Generator().manual_seed(7)).cuda()
        (y * cot).sum().backward()
        named = dict(mod.named_parameters())
        for k, n in case["grad_norms"].items():
            if n > 1e-4:
                got = float(named[k].grad.double().norm())
                assert abs(got - n) / n < 5e-3, (k, got, n)
    finally:
        ab.set_precision("bf16")


def _cfg2_model(cls=28, T=32, V=22, seed=5):
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    st = O.random_state(O.model_spec(3, cls, T, V), seed)
    mod = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style="ST", graph="graph.SHRE", graph_args={"labeling_mode": "spatial"})
    mod.load_state_dict(st)
    return mod.cuda()


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_full_size_eval_is_batch_split_invariant(mode, tol):
    """BASELINE cfg2 size (256 sequences): in eval mode every sequence is independent, so the logits of the full
    batch must equal the logits of its 32-sequence slices -- a size-independent property checked where the oracle
    is too slow to run (tile walks, persistent-CTA schedules and split reductions all change with the batch).
    fp32 mode: to the parity tolerance.  bf16 mode: to bf16 noise only -- gcn0 centres its operands by the batch
    mean before rounding them to bf16 (exact in real arithmetic), so the rounding pattern depends on the batch."""
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    ab.set_precision(mode)
    try:
        mod = _cfg2_model().eval()
        x, _ = O.synthetic_batch(256, 32, 22, 28, 99)
        x = x.cuda()
        with torch.no_grad():
            full = mod(x).float()
            parts = torch.cat([mod(x[i:i + 32]).float() for i in range(0, 256, 32)])
        assert torch.isfinite(full).all()
        err = float((full - parts).norm() / full.norm())
        assert err < tol, err
    finally:
        ab.set_precision("bf16")


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-4), ("bf16", 5e-2)])
def test_full_size_gradient_is_additive_over_shards(mode, tol):
    """cfg2 size: with BatchNorm on running statistics and DropPath off the loss is a plain sum over sequences, so the
    gradient of the full batch equals the sum of the gradients of its two halves -- linearity of the backward kernels
    (dW split reductions, attention / LayerNorm backward, conv taps) at full size.  fp32 mode checks the kernels'
    arithmetic; bf16 mode can only agree to the noise of independently rounded activations / gradients."""
    import altformer_b200 as ab
    from altformer_b200 import functional as AF
    from oracle import altformer_oracle as O
    ab.set_precision(mode)
    try:
        mod = _cfg2_model().eval()            # running-stat BN and no DropPath
        for prm in mod.gcn0.parameters():     # gcn0's backward exists for batch-statistics BatchNorm only (training)
            prm.requires_grad_(False)
        x, y = O.synthetic_batch(256, 32, 22, 28, 123)
        x, y = x.cuda(), y.cuda()

        def grads(xs, ys, scale):
            mod.zero_grad(set_to_none=True)
            loss = AF.cross_entropy(mod(xs), ys) * scale   # cross_entropy averages: undo it so the pieces add up
            loss.backward()
            return torch.cat([p.grad.flatten().double() for _, p in mod.live_parameters() if p.grad is not None])

        g_full = grads(x, y, 256.0)
        g_a, g_b = grads(x[:128], y[:128], 128.0), grads(x[128:], y[128:], 128.0)
        err = float((g_full - (g_a + g_b)).norm() / g_full.norm())
        assert torch.isfinite(g_full).all() and err < tol, err
    finally:
        ab.set_precision("bf16")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_cfg2_train_step_against_oracle(mode):
    """The benchmarked configuration itself (BASELINE configs[1]: 256 sequences, T=32, V=22, 28 classes, style ST, train-mode
    BatchNorm, cross-entropy) against the CPU oracle: logits, loss and EVERY live gradient tensor.
    fp32 mode: logits 5e-5 (north_star: 1e-4); gradients of everything downstream of the max-over-T pool 1e-4; the other
    tensors median 3e-3 / worst 3e-2 -- at this size a handful of the 131,072 arg-max decisions of the pool are near-ties
    within the 1e-5 kernel error and flip, which re-routes gradient entries (tools/argmax_sensitivity.py reproduces the very
    same per-tensor pattern inside the fp64 oracle with a 1e-5 perturbation; full-size Blocks alone are at 1e-5, group
    `modules`).  bf16 mode: logits <= max(1e-2, 1.5 x) and every gradient tensor within tools.gpu_diag_modules.floor_bound of
    the error the reference math itself shows under torch.autocast(bf16) on this very case (tests/golden/autocast_floor.json,
    key cfg2_ST_N256), median within 2 x; the oracle quantises the gcn0 -> tcn0 activation to bf16 as the module boundary
    stores it (O.boundary_bf16)."""
    import altformer_b200 as ab
    from altformer_b200 import functional as AF
    from oracle import altformer_oracle as O
    from tools.gpu_diag_modules import autocast_floor, floor_bound, is_zero_class, noisy_class_rms
    N, T, V, cls = 256, 32, 22, 28
    A = O.spatial_graph(V)
    st = O.random_state(O.model_spec(3, cls, T, V), 5)
    x, labels = O.synthetic_batch(N, T, V, cls, 123)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in st.items()}
    torch.set_num_threads(max(len(__import__("os").sched_getaffinity(0)), 1))
    yr = O.model_forward(x, params, A, "ST", True, boundary=O.boundary_bf16 if mode == "bf16" else None)
    loss_r = torch.nn.functional.cross_entropy(yr, labels)
    loss_r.backward()
    ab.set_precision(mode)
    try:
        mod = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style="ST", graph="graph.SHRE", graph_args={"labeling_mode": "spatial"})
        mod.load_state_dict(st)
        mod = mod.cuda().train()
        for m in mod.modules():
            if type(m).__name__ == "DropPath":
                m.drop_prob = 0.0
        y = mod(x.cuda())
        loss = AF.cross_entropy(y, labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-300))  # noqa: E731
        e_logits = rel(y, yr.detach())
        floor = autocast_floor("cfg2_ST_N256") if mode == "bf16" else None
        rows = []
        for k, p in mod.named_parameters():
            rg = params[k].grad if k in params else None
            if rg is None or float(rg.norm()) == 0.0 or is_zero_class(k, True):
                continue
            assert p.grad is not None, k
            e = rel(p.grad, rg)
            bound = 3e-2 if mode == "fp32" else (floor_bound(floor, k) if floor else 1e-1)
            rows.append((e / bound, e, bound, k))
        rows.sort(reverse=True)
        es = sorted(r[1] for r in rows)
        print(f"cfg2 {mode}: logits rel_l2 {e_logits:.3e}, loss {float(loss):.6f} vs {float(loss_r):.6f}; {len(rows)} gradient tensors: median "
              f"{es[len(es) // 2]:.3e}, worst {es[-1]:.3e}; closest to its bound: {rows[0][3]} {rows[0][1]:.3e} / {rows[0][2]:.3e}")
        try:   # per-tensor table for profiles/ (best effort: the directory exists on the GPU box)
            import os
            os.makedirs("gpurun_out", exist_ok=True)
            with open(f"gpurun_out/cfg2_{mode}_grad_errors.txt", "w") as f:
                f.write(f"# cfg2 (N=256 T=32 V=22 28 classes, style ST, train) {mode} mode vs CPU oracle; logits rel_l2 {e_logits:.3e}\n")
                f.write("# tensor  rel_l2  bound" + ("  reference_autocast_bf16" if floor else "") + "\n")
                for k, p in mod.named_parameters():
                    hit = [r for r in rows if r[3] == k]
                    if hit:
                        f.write(f"{k:52s} {hit[0][1]:.3e} {hit[0][2]:.3e}" + (f" {floor['grads'].get(k, 0.0):.3e}" if floor else "") + "\n")
        except OSError:
            pass
        assert len(rows) >= 170
        if mode == "fp32":
            assert e_logits < 5e-5 and abs(float(loss) - float(loss_r)) < 1e-4
            assert es[len(es) // 2] < 3e-3, es[len(es) // 2]
            post_pool = [(k, e) for _, e, _, k in rows if "mlp_head" in k or k.endswith("modelA.blocks.5.mlp.fc2.bias")]
            assert len(post_pool) == 5 and all(e < 1e-4 for _, e in post_pool), post_pool
        else:
            assert e_logits < max(1e-2, 1.5 * (floor["logits"] if floor else 1e-2)) and abs(float(loss) - float(loss_r)) < 2e-2
            if floor:   # the cancellation-dominated tensors as a class: within 2 x the reference's own autocast error
                ours_rms, floor_rms = noisy_class_rms(rows, floor)
                print(f"cfg2 bf16: cancellation-dominated class RMS error {ours_rms:.3e} (reference autocast {floor_rms:.3e})")
                assert ours_rms <= 2.0 * floor_rms, (ours_rms, floor_rms)
            if floor:
                fl = sorted(floor["grads"].get(r[3], 0.0) for r in rows)
                print(f"     reference under autocast(bf16): median {fl[len(fl) // 2]:.3e}, worst {fl[-1]:.3e}")
                assert es[len(es) // 2] <= max(1e-2, 2.0 * fl[len(fl) // 2]), (es[len(es) // 2], fl[len(fl) // 2])
        bad = [(k, e, b) for _, e, b, k in rows if not e <= b]
        assert not bad, bad[:8]
    finally:
        ab.set_precision("bf16")


def test_ddp_world2_against_oracle():
    """Two ranks over NCCL (skipped with fewer than 2 GPUs): the all-reduced mean gradient equals the oracle run on every shard
    separately and averaged, ranks end bit-identical although they were built from different seeds, and unequal
    shards reduce to the global-batch mean.  tools/gpu_ddp_check.py holds the check; this test launches it under torchrun."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "tools", "gpu_ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and "ddp-check: PASS" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_reference_trained_checkpoint_loads_and_matches():
    """a DataParallel checkpoint as the reference's scripts write it ('module.'-prefixed keys,
    SHREC/ST_TS/emsemble.py:99-104) loads into the CUDA model and reproduces the reference's eval logits -- which requires
    the reference's EFFECTIVE adjacency (1e-6, model/unit_agcn.py:36-38,75), switched on by load_state_dict for such files."""
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    from tests import goldenlib as G
    case = G.load("model_ST_22_refckpt")
    N, T, V, cls = case["shape"]
    st = O.random_state(O.model_spec(3, cls, T, V), case["state_seed"])
    ckpt = {"module." + k: v for k, v in st.items()}
    x, _ = O.synthetic_batch(N, T, V, cls, case["batch_seed"])
    ab.set_precision("fp32")
    try:
        mod = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style="ST", graph="graph.SHRE", graph_args={"labeling_mode": "spatial"}).cuda()
        res = mod.load_state_dict(ckpt)          # after .cuda(): the adjacency switch must act on the device buffer
        assert not res.missing_keys and not res.unexpected_keys
        assert float(mod.gcn0.A.max()) == pytest.approx(1e-6)
        mod.eval()
        with torch.no_grad():
            y = mod(x.cuda())
        G.check_entry(case["y"], y, 2e-4, "eval logits of a reference-trained checkpoint")
        # the same tensors without the prefix are a checkpoint of THIS package: graph adjacency, different logits
        mod.load_state_dict(st)
        assert float(mod.gcn0.A.max()) > 1e-3
        with torch.no_grad():
            y2 = mod(x.cuda())
        assert float((y2.cpu() - case["y"]).norm() / case["y"].norm()) > 1e-2
        ab.set_precision("bf16")
        mod.load_state_dict(ckpt)
        with torch.no_grad():
            yb = mod(x.cuda())
        G.check_entry(case["y"], yb.float(), 2e-2, "bf16 eval logits of a reference-trained checkpoint")
    finally:
        ab.set_precision("bf16")


def test_streams_against_reference_hand_dataset_gpu():
    """Device stream transforms vs the reference's own Hand_Dataset.motion / bone / palm normalisation outputs."""
    import altformer_b200 as ab
    from tests import goldenlib as G
    case = G.load("streams_22")
    x = case["x"][None].cuda()
    palm = ab.streams.palm_normalise(x)
    assert torch.allclose(palm[0].cpu(), case["palm"], atol=1e-6)
    assert torch.allclose(ab.streams.motion(palm)[0].cpu(), case["motion"], atol=1e-6)
    assert torch.allclose(ab.streams.bone(palm)[0].cpu(), case["bone"], atol=1e-6)


@pytest.mark.gpu
def test_augment_against_reference_data_aug_gpu():
    """Hand_Dataset.data_aug on the device (afb_augment) vs the reference's own outputs for recorded draws
    (augment_22.pt) and vs the oracle on a larger seeded batch; random_augment draws stay inside the reference's ranges."""
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    from tests import goldenlib as G
    case = G.load("augment_22")
    y = ab.streams.augment(case["x"].cuda(), case["kind"].cuda(), case["params"].cuda())
    assert torch.allclose(y.cpu(), case["y"], atol=1e-6)
    g = torch.Generator().manual_seed(3)
    N, T, V = 67, 32, 22
    x = torch.randn(N, T, V, 3, generator=g)
    kind = torch.randint(-1, 5, (N,), generator=g, dtype=torch.int32)          # incl. "no transform" values
    params = torch.rand(N, 16, generator=g)
    params[:, :4] = torch.rand(N, V, generator=g).argsort(1)[:, :4].float()    # distinct joints for the noise transform
    got = ab.streams.augment(x.cuda(), kind.cuda(), params.cuda()).cpu()
    p_noise = params.clone()
    want = O.augment(x, kind, torch.where((kind == 2)[:, None], p_noise, params))
    # kinds other than noise read params[:, :3] as factor / offset / r -- the joint ids written above are just numbers there
    assert torch.allclose(got, want, atol=1e-6)
    gen = torch.Generator(device="cuda").manual_seed(9)
    xa, k, p = ab.streams.random_augment(x.cuda(), generator=gen)
    assert xa.shape == x.shape and set(k.cpu().tolist()) <= {0, 1, 2, 3} and len(set(k.cpu().tolist())) == 4
    pc, kc = p.cpu(), k.cpu()
    assert ((pc[kc == 0, 0] >= 0.8) & (pc[kc == 0, 0] <= 1.2)).all() and (pc[kc == 1, :3].abs() <= 0.1).all()
    assert (pc[kc == 2, 4:].abs() <= 0.1).all() and ((pc[kc == 3, 0] >= 0) & (pc[kc == 3, 0] <= 1)).all()
    j = pc[kc == 2, :4]
    assert ((j >= 0) & (j < V)).all() and all(len(set(r.tolist())) == 4 for r in j)
    assert torch.allclose(xa.cpu(), O.augment(x, kc, pc), atol=1e-6)


@pytest.mark.parametrize("style", ["STR", "TTR"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_strttr_against_reference_golden(style, mode):
    """the STR / TTR ablation models (altformer_b200.STR_TTR, drop-in for STR_TTR/STR_TTR.py) on the CUDA kernels
    vs the unmodified reference's train-mode output and parameter gradients (strttr_*_22.pt); strict state_dict load."""
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    from tests import goldenlib as G
    from tools.gpu_diag_modules import is_zero_class
    d = G.load(f"strttr_{style}_22")
    N, T, V, cls = d["shape"]
    st = O.random_state(O.strttr_spec(style, 3, cls, T, V), d["state_seed"])
    ab.set_precision(mode)
    try:
        m = ab.STR_TTR(3, cls, num_frame=T, num_joints=V, style=style, graph="graph.SHRE", graph_args={"labeling_mode": "spatial"})
        res = m.load_state_dict(st, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        for mod in m.modules():
            if type(mod).__name__ == "DropPath":
                mod.drop_prob = 0.0
        m = m.cuda().train()
        x, _ = O.synthetic_batch(N, T, V, cls, d["batch_seed"])
        y = m(x.cuda())
        assert y.dtype == torch.float32 and tuple(y.shape) == tuple(d["y"].shape)
        tol_y, tol_g = (1e-4, 2e-3) if mode == "fp32" else (2e-2, 1.5e-1)
        G.check_entry(d["y"], y, tol_y, f"{style} output")
        (y * d["w"].cuda()).sum().backward()
        errs = []
        for k, p in m.named_parameters():
            g = d["grads"].get(k)
            if g is None or p.grad is None:
                continue
            ref_norm = float(g.double().norm()) if torch.is_tensor(g) else g["norm"]
            if ref_norm < 1e-6 or is_zero_class(k, True):   # biases in front of a batch-stat BatchNorm: exact gradient 0
                continue
            got = p.grad.detach().cpu()
            if torch.is_tensor(g):
                errs.append((float((got.double() - g.double()).norm() / g.double().norm()), k))
            else:
                flat = got.reshape(-1)
                errs.append((float((flat[::g["stride"]].double() - g["sample"].double()).norm() / g["sample"].double().norm()), k))
        errs.sort()
        assert len(errs) > 80
        print(f"{style} {mode}: output ok, {len(errs)} gradient tensors, median {errs[len(errs) // 2][0]:.3e}, worst {errs[-1][0]:.3e} ({errs[-1][1]})")
        assert errs[len(errs) // 2][0] <= (2e-4 if mode == "fp32" else 3e-2), errs[len(errs) // 2]
        # bf16: the theta / phi convolutions and PA are cancellation-dominated (tools.gpu_diag_modules.floor_bound): their
        # relative error moves 0.04 ... 0.16 from run to run at this size; they only have to keep the order of magnitude
        noisy = ("conv_a.", "conv_b.", ".PA", "pos_embed")
        bad = [(e, k) for e, k in errs if e > (1.0 if mode == "bf16" and any(t in k for t in noisy) else tol_g)]
        assert not bad, bad[-3:]
    finally:
        ab.set_precision("bf16")


@pytest.mark.parametrize("name", ["unit2d_dim3_train", "unit2d_dim3_s2_train"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_unit2d_dim3_against_reference_golden(name, mode):
    """Unit2D(dim=3) (convolution along the joints, model/net.py:29-36), stride 1 and 2, train mode: output, input
    gradient, parameter gradients, running statistics.  fp32 mode vs the reference class's golden (1e-4); bf16 mode vs the
    oracle (pinned by the same golden) on the bf16-rounded input the module actually consumes, rel-L2 1e-2 / 2e-2 -- an
    input that is not bf16-representable moves ReLU masks, which is the input's rounding, not the kernels'."""
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    from tests import goldenlib as G
    case = G.load(name)
    cin, cout, k, N, T, V, stride = case["shape"]
    spec = O.unit2d_spec("", cin, cout, k)
    spec["conv.weight"] = (cout, cin, 1, k)
    st = O.random_state(spec, case["seed"])
    x0 = torch.randn(N, cin, T, V, generator=torch.Generator().manual_seed(case["seed"] + 100))
    cot = torch.randn(case["y"].shape, generator=torch.Generator().manual_seed(7))
    if mode == "bf16":
        x0 = x0.bfloat16().float()
        p = {kk: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in kk else v.clone()) for kk, v in st.items()}
        xr = x0.clone().requires_grad_(True)
        yr = O.unit2d_forward(xr, p, "", True, stride)
        (yr * cot).sum().backward()
        want = {"y": yr.detach(), "dx": xr.grad, "grad.conv.weight": p["conv.weight"].grad, "grad.bn.weight": p["bn.weight"].grad,
                "grad.bn.bias": p["bn.bias"].grad, "buf.bn.running_mean": p["bn.running_mean"], "buf.bn.running_var": p["bn.running_var"]}
    else:
        want = case
    ab.set_precision(mode)
    try:
        m = ab.Unit2D(cin, cout, kernel_size=k, stride=stride, dim=3)
        m.load_state_dict(st, strict=True)
        m = m.cuda().train()
        x = x0.cuda().requires_grad_(True)
        y = m(x)
        (y.float() * cot.cuda()).sum().backward()
        got = {"y": y.float(), "dx": x.grad, "grad.conv.weight": m.conv.weight.grad, "grad.bn.weight": m.bn.weight.grad,
               "grad.bn.bias": m.bn.bias.grad, "buf.bn.running_mean": m.bn.running_mean, "buf.bn.running_var": m.bn.running_var}
        for key, g in got.items():
            if mode == "fp32":
                G.check_entry(want[key], g, 1e-3 if key.startswith("buf.") else 1e-4, key)
            else:
                e = float((g.detach().cpu().double() - want[key].double()).norm() / want[key].double().norm())
                assert e <= (1e-2 if key in ("y", "buf.bn.running_mean", "buf.bn.running_var") else 2e-2), (key, e)
    finally:
        ab.set_precision("bf16")


def test_graphed_inference_matches_eager_eval():
    """ab.GraphedInference (CUDA-graph replay of the eval forward) returns the eager eval logits (to the 1-ulp run-to-run
    jitter the eager forward itself has from its atomically-accumulated reductions), for device and pinned-host inputs, and
    rejects another shape."""
    import altformer_b200 as ab
    from oracle import altformer_oracle as O
    N, T, V, cls = 8, 16, 22, 14
    torch.manual_seed(0)
    m = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style="ST", graph="graph.SHRE", graph_args={"labeling_mode": "spatial"}).cuda().eval()
    x1, _ = O.synthetic_batch(N, T, V, cls, 5)
    x2, _ = O.synthetic_batch(N, T, V, cls, 6)
    with torch.no_grad():
        want1, want2 = m(x1.cuda()).clone(), m(x2.cuda()).clone()
    g = ab.GraphedInference(m, x1.cuda())
    assert torch.allclose(g(x1.cuda()).clone(), want1, atol=1e-6, rtol=0)
    assert torch.allclose(g(x2.pin_memory()).clone(), want2, atol=1e-6, rtol=0)
    assert torch.allclose(g(x1.cuda()), want1, atol=1e-6, rtol=0)
    assert not torch.allclose(want1, want2, atol=1e-3)
    with pytest.raises(RuntimeError):
        g(x1[:4].cuda())
