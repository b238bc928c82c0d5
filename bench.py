"""Benchmark of the north-star hot path: SHREC'17-shape ST_GCN_AltFormer training step (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference] [--config 1|2|3|4|5]

--config 2 (default; the line the driver records): one "step" = zero_grad -> forward -> cross-entropy -> backward ->
gradient all-reduce (N > 1) -> AdamW on one batch of synthetic skeletons (batch 256 per GPU, T=32, V=22, 28 classes, bf16
compute with fp32 master weights / accumulation, style 'ST').  Weak scaling: every rank processes its own 256-sample shard;
gradients are averaged with one NCCL all-reduce.  With N > 1 the same line also carries `strong_scaling`: the SURVEY 8(e)
split of ONE 256-sample global batch into 256/N per GPU.

The other BASELINE configs are separate invocations with the same one-line schema (their JSON is kept under profiles/):
  --config 1   SHREC-shape forward, batch 32, eval, 14 classes (the reference's CPU-runnable case) on the GPU
  --config 3   unit_agcn + temporal-conv stack sweep, C = 64/128/256, T = 32/64/128, V = 22, global batch 1024, fwd + bwd
  --config 4   LMDHG shape (46 joints, T = 64, 14 classes), 64 sequences per GPU (= 512 over 8 GPUs), full training step
  --config 5   inference ensemble: joint + bone + motion streams x (ST, TS) models, batch 8192 per GPU

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the public API with
pinned-host inputs (H2D inside the timed region) and a D2H read of the result every step.  `roofline` is the gcn0
(unit_agcn 3->128) forward, the HBM-bound kernel the metric names; `roofline_tensor` is the whole step against the measured
dense bf16 peak.  `cpu_baseline` (N = 1 only) times the CPU reference path (the imported reference modules when
/root/reference is present, else the oracle port); `eager_gpu_baseline` (N = 1, config 2) times the same plain-PyTorch
restatement on the B200 (fp32 with TF32 off, and under autocast(bf16)) -- the informative same-box neighbour.
`--impl reference` prints the CPU arm as its own line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    1: dict(T=32, V=22, cls=14, per_gpu_batch=32, style="ST", graph="graph.SHRE", kind="infer",
            metric="AltFormer forward seq/s (SHREC shape, batch 32, eval)",
            workload="configs[0]: SHREC'17-shape ST_GCN_AltFormer(style ST) forward, batch 32, eval, T=32 V=22 14 classes"),
    2: dict(T=32, V=22, cls=28, per_gpu_batch=256, style="ST", graph="graph.SHRE", kind="train",
            metric="AltFormer train seq/s (SHREC shape)",
            workload="configs[1]: SHREC'17-shape ST_GCN_AltFormer(style ST) fwd+bwd+AdamW training step, T=32 V=22 28 classes"),
    3: dict(T=32, V=22, cls=0, per_gpu_batch=1024, style=None, graph="graph.SHRE", kind="stack",
            metric="unit_agcn+Unit2D stack fwd+bwd seq/s (DHG shape)",
            workload="configs[2]: 4 x TCN_GCN_unit(C, C) (unit_agcn -> Unit2D 9x1 -> +x), V=22, global batch 1024, fwd+bwd, sweep over C and T"),
    4: dict(T=64, V=46, cls=14, per_gpu_batch=64, style="TS", graph="graph.LMDHG", kind="train",
            metric="AltFormer train seq/s (LMDHG shape)",
            workload="configs[3]: LMDHG-shape (46 joints) ST_GCN_AltFormer fwd+bwd+AdamW training step, T=64 14 classes, 64 sequences per GPU (512 over 8 GPUs)"),
    5: dict(T=32, V=22, cls=28, per_gpu_batch=8192, style="both", graph="graph.SHRE", kind="ensemble",
            metric="AltFormer ensemble inference seq/s (SHREC shape)",
            workload="configs[4]: joint + bone + motion streams x (ST, TS) models, 0.8 ST + 0.2 TS, summed over streams, eval, batch 8192 per GPU"),
}
CFG = dict(CONFIGS[2])


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def flops_per_sample_fwd(T, V, cls, style="ST"):
    """transformer block 16 D^2 M + 4 L D M, embeds, tcn0, gcn0, head."""
    blk = lambda D, M, L: 16 * D * D * M + 4 * L * D * M  # noqa: E731
    if style == "TS":
        M1, M2, L1, L2 = T * V, V, T, V
    else:
        M1, M2, L1, L2 = T * V, T, V, T
    f = 6 * blk(256, M1, L1) + 6 * blk(512, M2, L2)
    f += 2 * 128 * 256 * M1 + 2 * 256 * 512 * M2 + 2 * 512 * cls
    f += 2 * 128 * 128 * 9 * T * V + 7.13e6 * (T * V) / (32 * 22)
    return f


def synthetic_batch(N, T, V, cls, seed=1234):
    """BASELINE.md 4 / SURVEY 8d inputs: 0.2*randn skeletons, palm-centred (Hand_Dataset.py:61), uniform labels.
    (Own copy: the measured arm does not import anything from oracle/.)"""
    g = torch.Generator().manual_seed(seed)
    x = 0.2 * torch.randn(N, T, V, 3, generator=g)
    x = x - x[:, :1, 1:2, :]
    y = torch.randint(0, max(cls, 1), (N,), generator=g)
    return x, y


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline is meant to use every core it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return n


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores (imported reference modules when present, else the oracle port)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_model(cfg, style, seed=0, train=True):
    """(model, kind): the unmodified reference module tree if /root/reference (or $ALTFORMER_REFERENCE) exists -- it does
    in the authoring container, not on the GPU box -- else the oracle port.  Same seeded weights either way."""
    from oracle import altformer_oracle as O
    from oracle import refshim
    T, V, cls = cfg["T"], cfg["V"], cfg["cls"]
    state = O.random_state(O.model_spec(3, cls, T, V), seed)
    if refshim.available() and hasattr(refshim, "build_model"):
        try:
            return refshim.build_model(state, cls, T, V, style, cfg["graph"], train), "reference"
        except Exception as e:  # noqa: BLE001  (the port is the documented fallback)
            print(f"[bench] reference modules unusable ({e!r}); using the oracle port", file=sys.stderr)
    m = O.OracleModel(state, O.spatial_graph(V), style)
    return m.train(train), "port"


def cpu_train_leg(cfg, B=32, budget_s=20.0, max_steps=12, min_steps=2):
    """Bounded sample of the train step on the CPU: fwd + CE + zero_grad + bwd + AdamW at batch B."""
    from oracle import altformer_oracle as O
    cores = use_all_host_threads()
    model, kind = _cpu_model(cfg, cfg["style"])
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.1)
    x, y = O.synthetic_batch(B, cfg["T"], cfg["V"], cfg["cls"])
    crit = torch.nn.CrossEntropyLoss()
    times, t_start = [], time.perf_counter()
    while len(times) < min_steps + 1 or (time.perf_counter() - t_start < budget_s and len(times) < max_steps):
        t0 = time.perf_counter()
        loss = crit(model(x), y)
        model.zero_grad()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(times[1:])
    cores = torch.get_num_threads()
    return {"value": B / (ms / 1e3), "unit": "seq/s", "cores": cores, "kind": kind, "ms_per_step": ms,
            "sample": f"{len(times) - 1} train steps of batch {B} (of the {cfg['per_gpu_batch']}-sample step), fp32, {cores} threads"}


def cpu_infer_leg(cfg, B=32, budget_s=15.0, styles=("ST",), max_steps=12):
    """Eval-mode forward on the CPU (configs[0] exactly when cfg is CONFIGS[1]: batch 32, 14 classes, fp32)."""
    from oracle import altformer_oracle as O
    use_all_host_threads()
    models, kind = [], "port"
    for i, s in enumerate(styles):
        m, kind = _cpu_model(cfg, s, seed=i, train=False)
        models.append(m)
    x, _ = O.synthetic_batch(B, cfg["T"], cfg["V"], cfg["cls"])
    times, t_start = [], time.perf_counter()
    with torch.no_grad():
        while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < max_steps):
            t0 = time.perf_counter()
            for m in models:
                m(x)
            times.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(times[1:])
    cores = torch.get_num_threads()
    return {"value": B / (ms / 1e3), "unit": "seq/s", "cores": cores, "kind": kind, "ms_per_step": ms,
            "sample": f"{len(times) - 1} eval forwards of batch {B} through {len(models)} model(s), fp32, {cores} threads"}


def cpu_stack_leg(C, T, V, layers, B=8, budget_s=8.0):
    """fwd + bwd of the TCN_GCN_unit stack on the CPU (oracle port; the reference's TCN_GCN_unit needs its whole ST_TR file)."""
    from oracle import altformer_oracle as O
    use_all_host_threads()
    A = O.spatial_graph(V)
    spec = O.OrderedDict()
    for l in range(layers):
        spec.update(O.agcn_spec(f"l{l}.gcn1.", C, C, V))
        spec.update(O.unit2d_spec(f"l{l}.tcn1.", C, C, 9))
    st = O.random_state(spec, 3)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in st.items()}
    x = torch.randn(B, C, T, V)
    times, t_start = [], time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 8):
        t0 = time.perf_counter()
        h = x
        for l in range(layers):
            h = O.tcn_gcn_forward(h, params, f"l{l}.", A, True)
        h.square().mean().backward()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(times[1:])
    return {"value": B / (ms / 1e3), "unit": "seq/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times) - 1} fwd+bwd passes of batch {B}, C={C} T={T}, fp32"}


def cpu_leg(cfg, args=None):
    if cfg["kind"] == "train":
        B = 32 if cfg["V"] == 22 else 8
        return cpu_train_leg(cfg, B=B)
    if cfg["kind"] == "infer":
        return cpu_infer_leg(cfg, B=32)
    if cfg["kind"] == "ensemble":
        leg = cpu_infer_leg(cfg, B=32, styles=("ST", "TS"), budget_s=20.0)
        # 3 streams x (ST, TS) = 3 x the timed pair; the stream transforms themselves are negligible
        leg["value"] /= 3.0
        leg["sample"] += "; one (ST, TS) pair timed, x3 streams"
        return leg
    return cpu_stack_leg(128, 32, cfg["V"], 4)


def run_reference(args):
    """`--impl reference`: the CPU arm as its own JSON line (rank 0 only; other ranks exit without work)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = CFG
    leg = cpu_leg(cfg)
    emit({
        "impl": "reference", "metric": cfg["metric"], "value": leg["value"], "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": leg.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["workload"]},
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------------------------------------------------
# timing helpers
# ---------------------------------------------------------------------------------------------------------------------
def timed(fn, steps, dist_on, dev):
    """barrier + sync, K steps between CUDA events, sync + barrier; returns (max-over-ranks ms/step, median ms of the
    per-step event pairs on this rank)."""
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(dev)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record()
    for i in range(steps):
        fn()
        marks[i + 1].record()
    torch.cuda.synchronize(dev)
    ms = marks[0].elapsed_time(marks[-1]) / steps
    med = statistics.median(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
    if dist_on:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t)
    return ms, med


def gcn0_traffic():
    """dram bytes (read + write) per gcn0 forward launch from the committed ncu --set full summary of the shipped kernel."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_gcn0_final.json")
    try:
        d = json.load(open(path))
        # one cold launch under ncu leaves the 46 MB output in the 126 MB L2 (dram write ~0); the bytes the kernel stored
        # (l2_bytes_stored) are what reaches HBM when launches stream over distinct buffers, as the timed graph does
        wr = max(float(d["dram_bytes_write"]), float(d.get("l2_bytes_stored", 0.0)))
        return float(d["dram_bytes_read"]) + wr, os.path.relpath(path, ROOT)
    except (OSError, KeyError, ValueError):
        return None, None


def gcn0_roofline(model, x, dev, pk, sets=8, reps=7):
    """gcn0 forward alone.  One CUDA graph of `sets` launches over `sets` DISTINCT input/output buffer sets (sets x 48 MB
    of output > the 126 MB L2, so every launch streams to HBM and finds none of its lines cached), CUDA events around a
    replay, median over `reps` replays -> time per launch without host launch gaps."""
    N, T, V, _ = x.shape
    xs = [x] + [(x + 0.01 * (i + 1)).contiguous() for i in range(sets - 1)]
    with torch.no_grad():
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for xi in xs[:2]:
                model.gcn0.forward_skeleton(xi)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            keep = [model.gcn0.forward_skeleton(xi) for xi in xs]
        times = []
        for _ in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            times.append(e0.elapsed_time(e1) / sets)
    del keep
    ms = statistics.median(times[2:])
    alg_bytes = N * T * V * (3 * 4 + 128 * 2)
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    traffic, src = gcn0_traffic()
    return {"bound": "hbm", "kernel": "gcn0 = unit_agcn(3->128) forward: gcn0_fused_kernel (one cooperative launch: scores, M, z, moments, "
                                       "grid barrier, BN fold, expansion, TMA stores)",
            "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"], "frac_of_8TBps_nominal": achieved / 8000.0,
            "peak_source": pk["src"], "traffic": traffic if (N, T, V) == (256, 32, 22) else None, "traffic_source": src,
            "alg_bytes_per_launch": alg_bytes, "ms_per_launch": ms,
            "method": f"graph of {sets} launches over {sets} distinct buffer sets ({sets * alg_bytes / 1e6:.0f} MB > L2), median of {reps} replays"}


def step_kernel_rooflines(M, dev, pk, iters=10):
    """The kernels that carry the step's time (spatial stage, D = 256, M = per-GPU tokens), each timed live with CUDA
    events over `iters` back-to-back launches on tensors far larger than L2: algorithmic bytes / time vs the HBM peak."""
    from altformer_b200 import ops
    # the same launch repeated on the same tensors must not profit from the ping-pong traversal order (each repetition would
    # start on the lines the previous one left in L2, which no kernel of the real step does for its whole input)
    saved_pp = os.environ.get("AFB_PINGPONG")
    os.environ["AFB_PINGPONG"] = "0"
    mk = lambda r, c: (0.5 * torch.randn(r, c, device=dev)).to(torch.bfloat16)  # noqa: E731
    x256, x512, g768, res = mk(M, 256), mk(M, 512), mk(M, 768), mk(M, 256)
    wqkv, wfc2 = mk(768, 256), mk(256, 512)
    bq, b256 = torch.zeros(768, device=dev), torch.zeros(256, device=dev)
    dW = torch.zeros(768, 256, device=dev)
    gam, bet = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    B, L, H = M // 22, 22, 8
    cases = [
        ("gemm_tn qkv 768x256 +bias", lambda: ops.gemm_tn(x256, wqkv, 768, bias=bq), M * (256 + 768) * 2),
        ("gemm_tn fc2 256x512 +bias+residual", lambda: ops.gemm_tn(x512, wfc2, 256, bias=b256, residual=res), M * (512 + 512) * 2),
        ("gemm_dw qkv 768x256 +dbias", lambda: ops.gemm_dw(g768, x256, dW, dbias=bq), M * 1024 * 2),
        ("attention_fwd L=22 8x32", lambda: ops.attention_fwd(g768, B, L, H), M * 4 * 256 * 2),
        ("attention_bwd L=22 8x32", lambda: ops.attention_bwd(g768, x256, B, L, H), M * 7 * 256 * 2),
        ("layernorm_fwd 256", lambda: ops.layernorm_fwd(x256, gam, bet, 1e-6), M * 256 * 2 * 2),
    ]
    out = []
    for name, fn, nbytes in cases:
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        us = 1e3 * e0.elapsed_time(e1) / iters
        gbs = nbytes / us / 1e3
        out.append({"kernel": name, "us": round(us, 1), "achieved": round(gbs, 1), "unit": "GB/s", "frac": round(gbs / pk["hbm"], 3)})
    if saved_pp is None:
        os.environ.pop("AFB_PINGPONG", None)
    else:
        os.environ["AFB_PINGPONG"] = saved_pp
    return out


def eager_gpu_baseline(cfg, dev, B, steps=5):
    """The reference math as plain PyTorch (the oracle port: the same ATen calls the reference makes) on THIS GPU: full
    train step at batch B, fp32 with TF32 off and under autocast(bf16).  A baseline leg like cpu_baseline: it says what
    eager cuDNN/cuBLAS on sm_100 does with the same step (SURVEY 8d "the bar on the same box")."""
    from oracle import altformer_oracle as O
    out = {}
    x, y = O.synthetic_batch(B, cfg["T"], cfg["V"], cfg["cls"])
    x, y = x.to(dev), y.to(dev)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for name, ac in (("fp32_tf32_off", False), ("autocast_bf16", True)):
            model = O.OracleModel(O.random_state(O.model_spec(3, cfg["cls"], cfg["T"], cfg["V"]), 0), O.spatial_graph(cfg["V"]), cfg["style"]).to(dev).train()
            model.A = model.A.to(dev)
            model._b = {k: v.to(dev) for k, v in model._b.items()}
            opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.1)
            crit = torch.nn.CrossEntropyLoss()

            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    loss = crit(model(x).float(), y)
                model.zero_grad()
                loss.backward()
                opt.step()

            for _ in range(3):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": B / (ms * 1e-3), "unit": "seq/s", "ms_per_step": ms}
            del model, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    out["kind"] = "port"
    out["sample"] = f"{steps} train steps of batch {B} (fwd+CE+bwd+AdamW), plain PyTorch eager on this GPU, no CUDA graph"
    return out


_JSON_FD = None


def _reserve_stdout():
    """stdout must carry exactly ONE JSON line: libraries (NCCL's version banner, torchrun notices) also write to
    fd 1, so fd 1 is pointed at stderr for the run and the JSON line goes to a saved copy of the original."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


# ---------------------------------------------------------------------------------------------------------------------
# measured arm
# ---------------------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)
        torch.cuda.set_device(self.dev)
        self.dist_on = self.world > 1
        if self.dist_on:
            dist.init_process_group("nccl", device_id=self.dev)
        self.pk = peaks()
        self.args = args

    def close(self):
        import torch.distributed as dist
        if self.dist_on:
            dist.barrier()
            dist.destroy_process_group()


def build_model(cfg, dev, style=None, seed=0, cls=None):
    import altformer_b200 as ab
    torch.manual_seed(seed)
    model = ab.ST_GCN_AltFormer(3, cls or cfg["cls"], num_frame=cfg["T"], num_joints=cfg["V"], style=style if style is not None else cfg["style"],
                                graph=cfg["graph"], graph_args={"labeling_mode": "spatial"}).to(dev)
    with torch.no_grad():  # the reference's init makes gcn0 invisible (bn gamma 1e-6); use a live one for a fair workload
        model.gcn0.bn.weight.fill_(1.0)
    return model


def train_leg(env, cfg, B, steps, warmup, use_graph=True, seed_off=0):
    """(ms/step resident, median ms, ms/step e2e, launches, clocks, model, x_dev, bytes) for per-GPU batch B."""
    import altformer_b200 as ab
    from altformer_b200 import ops
    dev = env.dev
    model = build_model(cfg, dev)
    trainer = ab.DataParallelTrainer(model, use_graph=use_graph)
    x_cpu, y_cpu = synthetic_batch(B, cfg["T"], cfg["V"], cfg["cls"], 1234 + env.rank + seed_off)
    x_pin, y_pin = x_cpu.pin_memory(), y_cpu.pin_memory()
    x_dev, y_dev = x_pin.to(dev), y_pin.to(dev)

    def step_resident():
        trainer.step(x_dev, y_dev)

    # e2e: every step uploads its batch from pinned host memory (straight into the step's input buffers) and reads its
    # loss back (as get_acc does, train_sttran.py:105-109).  The read-back is asynchronous and consumed one step later,
    # so the host enqueues step i+1 while step i runs; the timed region ends with a full sync.
    host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    st = {"i": 0, "last": float("nan")}

    def step_e2e():
        k = st["i"] & 1
        loss, _ = trainer.step(x_pin, y_pin)
        host_loss[k].copy_(loss, non_blocking=True)
        loss_ready[k].record()
        if st["i"] > 0:
            loss_ready[k ^ 1].synchronize()
            st["last"] = float(host_loss[k ^ 1])
        st["i"] += 1

    for _ in range(max(warmup, 3)):
        step_resident()
    ops.LAUNCHES[0] = 0
    if not use_graph:
        step_resident()
        launches = ops.LAUNCHES[0]
    else:
        launches = trainer.launches_per_step
    with ClockSampler(env.local) as clk:
        ms, med = timed(step_resident, steps, env.dist_on, dev)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, steps, env.dist_on, dev)
    return dict(ms=ms, median=med, ms_e2e=ms_e2e, launches=launches, clocks=clk.summary(), model=model, x_dev=x_dev,
                h2d=x_pin.numel() * 4 + y_pin.numel() * 8, d2h=4, last_loss=st["last"])


def dropin_leg(cfg, dev, B, steps=10, warmup=3):
    """The drop-in path of INTEGRATION.md section 2 WITHOUT the trainer: the reference's own loop (train_sttran.py:89-102 --
    model(x), CrossEntropyLoss, loss.backward(), torch.optim.AdamW.step(), zero_grad) on the CUDA modules, stock autograd,
    no CUDA graph, no flat buffers.  Reported beside the trainer's step so the cost of staying on stock torch is on record."""
    from altformer_b200 import ops
    model = build_model(cfg, dev)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.1)
    x, y = synthetic_batch(B, cfg["T"], cfg["V"], cfg["cls"], 4321)
    x, y = x.to(dev), y.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    ops.LAUNCHES[0] = 0
    step()
    launches = ops.LAUNCHES[0]
    ms, med = timed(step, steps, False, dev)
    return {"value": B / (ms * 1e-3), "unit": "seq/s", "ms_per_step": ms, "ms_per_step_median": med, "own_kernel_launches": launches,
            "sample": f"{steps} steps of batch {B}: model(x) + cross_entropy + loss.backward() + torch.optim.AdamW, stock autograd, no trainer, no CUDA graph"}


def run_train(env, cfg, args):
    B = args.batch or cfg["per_gpu_batch"]
    leg = train_leg(env, cfg, B, args.steps, args.warmup, use_graph=not args.no_graph)
    strong = None
    if env.world > 1 and cfg is CONFIGS[2] and B % env.world == 0 and not args.no_strong:
        # ONE global batch of 256 split 256/N per GPU (the launch-bound, under-filled regime)
        Bs = B // env.world
        del leg["model"]
        torch.cuda.empty_cache()
        s = train_leg(env, cfg, Bs, args.steps, args.warmup, use_graph=not args.no_graph, seed_off=100)
        strong = {"global_batch": B, "per_gpu_batch": Bs, "ms_per_step": s["ms"], "value": B / (s["ms"] * 1e-3), "unit": "seq/s",
                  "e2e_value": B / (s["ms_e2e"] * 1e-3), "gpu_launches": s["launches"]}
        leg["model"], leg["x_dev"] = s["model"], None
    if env.rank != 0:
        return
    T, V, cls = cfg["T"], cfg["V"], cfg["cls"]
    gb = B * env.world
    ms, ms_e2e = leg["ms"], leg["ms_e2e"]
    flops_step = 3.0 * flops_per_sample_fwd(T, V, cls, cfg["style"]) * B
    tf = flops_step / (ms * 1e-3) / 1e12
    pk = env.pk
    out = {
        "metric": cfg["metric"], "value": gb / (ms * 1e-3), "unit": "seq/s", "n_gpus": env.world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "ms_per_step_median": leg["median"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["workload"] if not args.shape else f"ad-hoc --shape {args.shape}: ST_GCN_AltFormer fwd+bwd+AdamW training step",
                   "style": cfg["style"], "per_gpu_batch": B, "global_batch": gb, "parallelism": f"dp{env.world}", "cuda_graph": not args.no_graph,
                   "l2": "no explicit flush: the step streams >5 GB of activations per GPU, far beyond the 126 MB L2"},
        "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "seq/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": leg["h2d"], "d2h_bytes_per_step": leg["d2h"]},
        "gpu_launches": leg["launches"],
        "clocks": leg["clocks"],
        "roofline_tensor": {"bound": "tensor", "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"],
                            "peak_source": pk["src"] + " sustained", "note": "whole step, 3 x forward FLOPs (SURVEY 8d), per GPU"},
    }
    if strong is not None:
        out["strong_scaling"] = strong
    if leg.get("x_dev") is not None and V <= 24 and V % 2 == 0:
        out["roofline"] = gcn0_roofline(leg["model"], leg["x_dev"], env.dev, pk)
        if V == 22 and not args.shape:
            # the kernels the step actually spends its time in (gcn0 above is the metric's kernel but <1 % of the step)
            out["roofline_step_kernels"] = step_kernel_rooflines(B * T * V, env.dev, pk)
    else:
        out["roofline"] = out["roofline_tensor"]
    if env.world == 1 and not args.no_cpu_baseline:
        # N = 1 only: under torchrun the other ranks would idle in a barrier while rank 0 runs the CPU leg
        if cfg is CONFIGS[2] and not args.no_eager:
            del leg
            torch.cuda.empty_cache()
            try:
                out["dropin_autograd"] = dropin_leg(cfg, env.dev, B)
            except Exception as e:  # noqa: BLE001
                out["dropin_autograd"] = {"unavailable": repr(e)[:200]}
            torch.cuda.empty_cache()
            try:
                out["eager_gpu_baseline"] = eager_gpu_baseline(cfg, env.dev, B)
            except Exception as e:  # noqa: BLE001
                out["eager_gpu_baseline"] = {"unavailable": repr(e)[:200]}
        c = cpu_leg(cfg)
        out["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(out)


def run_infer(env, cfg, args):
    """configs[0] on the GPU: eval forward at batch 32 (e2e = pinned-host batch in, logits out)."""
    import altformer_b200 as ab  # noqa: F401
    dev, B = env.dev, args.batch or cfg["per_gpu_batch"]
    model = build_model(cfg, dev).eval()
    x_cpu, _ = synthetic_batch(B, cfg["T"], cfg["V"], cfg["cls"], 1234 + env.rank)
    x_pin = x_cpu.pin_memory()
    x_dev = x_pin.to(dev)
    host_out = torch.empty((B, cfg["cls"]), dtype=torch.float32).pin_memory()

    use_graph = not args.no_graph
    from altformer_b200 import ops
    with torch.no_grad():
        model(x_dev)                   # first call: derived-weight casts, lazy inits
        ops.LAUNCHES[0] = 0
        model(x_dev)
    launches = ops.LAUNCHES[0]
    graphed = ab.GraphedInference(model, x_dev) if use_graph else None   # the public fixed-shape serving call

    def fwd_resident():
        if use_graph:
            graphed(x_dev)
        else:
            with torch.no_grad():
                model(x_dev)

    def fwd_e2e():
        if use_graph:
            host_out.copy_(graphed(x_pin), non_blocking=True)
        else:
            with torch.no_grad():
                y = model(x_pin.to(dev, non_blocking=True))
                host_out.copy_(y.float(), non_blocking=True)

    for _ in range(max(args.warmup, 3)):
        fwd_resident()
    with ClockSampler(env.local) as clk:
        ms, med = timed(fwd_resident, args.steps, env.dist_on, dev)
    ms_e2e, _ = timed(fwd_e2e, args.steps, env.dist_on, dev)
    if env.rank != 0:
        return
    gb = B * env.world
    tf = flops_per_sample_fwd(cfg["T"], cfg["V"], cfg["cls"], cfg["style"]) * B / (ms * 1e-3) / 1e12
    out = {"metric": cfg["metric"], "value": gb / (ms * 1e-3), "unit": "seq/s", "n_gpus": env.world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms, "ms_per_step_median": med, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic", "config": {"workload": cfg["workload"], "per_gpu_batch": B, "global_batch": gb, "cuda_graph": use_graph,
                                           "l2": "batch 32: the activations fit L2, nothing is flushed (a 12 MB working set cannot be made HBM-bound honestly)"},
           "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "seq/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": x_pin.numel() * 4,
                   "d2h_bytes_per_step": host_out.numel() * 4},
           "gpu_launches": launches, "clocks": clk.summary(),
           "roofline": {"bound": "tensor", "achieved": tf, "peak": env.pk["tf"], "unit": "TFLOP/s", "frac": tf / env.pk["tf"],
                        "note": "forward FLOPs of the batch; at batch 32 the pass is bound by the latency of ~95 dependent small kernels, not by the roofline"}}
    if env.world == 1 and not args.no_cpu_baseline:
        c = cpu_leg(cfg)
        out["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(out)


def run_ensemble(env, cfg, args):
    """configs[4]: three input streams (joint, bone, motion) x (ST, TS) models, combined on device (emsemble.py:215-225)."""
    import altformer_b200 as ab
    dev, B = env.dev, args.batch or cfg["per_gpu_batch"]
    models = {}
    for i, s in enumerate(("joint", "bone", "motion")):
        models[s] = (build_model(cfg, dev, "ST", 10 + i).eval(), build_model(cfg, dev, "TS", 20 + i).eval())
    x_cpu, _ = synthetic_batch(B, cfg["T"], cfg["V"], cfg["cls"], 7 + env.rank)
    x_pin = x_cpu.pin_memory()
    x_dev = x_pin.to(dev)
    host_out = torch.empty((B, cfg["cls"]), dtype=torch.float32).pin_memory()

    def resident():
        ab.streams.ensemble_forward(x_dev, models)

    def e2e():
        y = ab.streams.ensemble_forward(x_pin.to(dev, non_blocking=True), models)
        host_out.copy_(y.float(), non_blocking=True)

    steps, warmup = min(args.steps, 10), min(max(args.warmup, 3), 5)
    for _ in range(warmup):
        resident()
    from altformer_b200 import ops
    ops.LAUNCHES[0] = 0
    resident()
    launches = ops.LAUNCHES[0]
    with ClockSampler(env.local) as clk:
        ms, med = timed(resident, steps, env.dist_on, dev)
    ms_e2e, _ = timed(e2e, steps, env.dist_on, dev)
    if env.rank != 0:
        return
    gb = B * env.world
    f = 3 * (flops_per_sample_fwd(cfg["T"], cfg["V"], cfg["cls"], "ST") + flops_per_sample_fwd(cfg["T"], cfg["V"], cfg["cls"], "TS")) * B
    tf = f / (ms * 1e-3) / 1e12
    out = {"metric": cfg["metric"], "value": gb / (ms * 1e-3), "unit": "seq/s", "n_gpus": env.world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms, "ms_per_step_median": med, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic", "model_forwards_per_s": 6 * gb / (ms * 1e-3),
           "config": {"workload": cfg["workload"], "per_gpu_batch": B, "global_batch": gb, "cuda_graph": False,
                      "l2": "no explicit flush: one pass streams tens of GB of activations"},
           "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "seq/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": x_pin.numel() * 4,
                   "d2h_bytes_per_step": host_out.numel() * 4},
           "gpu_launches": launches, "clocks": clk.summary(), "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": env.pk["tf"], "unit": "TFLOP/s", "frac": tf / env.pk["tf"],
                        "note": "forward FLOPs of the six model passes per batch"}}
    if env.world == 1 and not args.no_cpu_baseline:
        c = cpu_leg(cfg)
        out["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(out)


def stack_flops_bytes(C, V):
    """SURVEY 8d per position, forward: unit_agcn(C->C) FLOPs (theta/phi 12 C IC + scores 6 V IC + aggregate 6 C V + conv_d
    6 C^2 + ~10 C) + Unit2D 2 C^2 9; algorithmic bytes 2 C e per module (read + write, e = 2)."""
    IC = C // 4
    agcn = 12 * C * IC + 6 * V * IC + 6 * C * V + 6 * C * C + 10 * C
    return agcn + 2 * C * C * 9, 2 * (2 * C * 2)


def run_stack(env, cfg, args):
    """configs[2]: 4 x TCN_GCN_unit(C, C) fwd + bwd, global batch 1024 sharded over the ranks; sweep C x T.  Reported with
    the exact-ReLU-mask forward (parity mode of the bf16 path, the default) and with it off (plain bf16 operands)."""
    import altformer_b200 as ab
    from altformer_b200 import functional as AF
    from altformer_b200 import ops
    dev, V, layers = env.dev, cfg["V"], 4
    NB = (args.batch or cfg["per_gpu_batch"]) // env.world
    A = torch.as_tensor(ab.import_class(cfg["graph"])(labeling_mode="spatial").A, dtype=torch.float32)
    sweep = []
    Cs = [int(c) for c in args.stack_c.split(",")]
    Ts = [int(t) for t in args.stack_t.split(",")]
    for C in Cs:
        for T in Ts:
            row = {"C": C, "T": T, "per_gpu_batch": NB}
            for exact in (True, False):
                AF.set_exact_bn_mask(exact)
                torch.manual_seed(0)
                units = [ab.TCN_GCN_unit(C, C, A, dropout=0.0).to(dev).train() for _ in range(layers)]
                x = torch.randn(NB * T * V, C, device=dev).to(torch.bfloat16).requires_grad_(True)
                dims = (NB, T, V)

                def fwd_bwd():
                    h = x
                    for u in units:
                        h = u.forward_tokens(h, dims)
                    h.backward(h.detach())
                    x.grad = None
                    for u in units:
                        for p in u.parameters():
                            p.grad = None

                for _ in range(2):
                    fwd_bwd()
                ops.LAUNCHES[0] = 0
                fwd_bwd()
                launches = ops.LAUNCHES[0]
                steps = max(3, min(args.steps, 10))
                ms, _ = timed(fwd_bwd, steps, env.dist_on, dev)
                fl, by = stack_flops_bytes(C, V)
                pos = NB * T * V
                t_hbm = 3 * layers * by * pos / (env.pk["hbm"] * 1e9)
                t_tc = 3 * layers * fl * pos / (env.pk["tf"] * 1e12)
                key = "exact_mask" if exact else "plain_bf16"
                row[key] = {"ms": ms, "seq_per_s": NB * env.world / (ms * 1e-3), "gpu_launches": launches,
                            "frac_hbm_roofline": t_hbm / (ms * 1e-3), "frac_tensor_roofline": t_tc / (ms * 1e-3),
                            "frac_roofline": max(t_hbm, t_tc) / (ms * 1e-3)}
                del units, x
                torch.cuda.empty_cache()
            AF.set_exact_bn_mask(True)
            sweep.append(row)
            if env.rank == 0:
                print(f"[cfg3] {row}", file=sys.stderr, flush=True)
    if env.rank != 0:
        return
    head = next((r for r in sweep if r["C"] == 128 and r["T"] == 32), sweep[0])
    out = {"metric": cfg["metric"], "value": head["exact_mask"]["seq_per_s"], "unit": "seq/s", "n_gpus": env.world, "steps": max(3, min(args.steps, 10)),
           "warmup": 3, "ms_per_step": head["exact_mask"]["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic", "config": {"workload": cfg["workload"], "headline_point": {"C": head["C"], "T": head["T"]}, "layers": layers,
                                           "global_batch": NB * env.world, "cuda_graph": False,
                                           "l2": "activations of every point exceed L2 (>= 184 MB per tensor); nothing flushed explicitly"},
           "sweep": sweep, "gpu_launches": head["exact_mask"]["gpu_launches"],
           "roofline": {"bound": "tensor" if head["exact_mask"]["frac_tensor_roofline"] > head["exact_mask"]["frac_hbm_roofline"] else "hbm",
                        "frac": head["exact_mask"]["frac_roofline"], "note": "max(bytes/BW, FLOPs/tensor peak) / measured time, fwd+bwd = 3 x fwd (SURVEY 8d)"}}
    if env.world == 1 and not args.no_cpu_baseline:
        c = cpu_stack_leg(head["C"], head["T"], V, layers)
        out["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(out)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)      # >= 20 warm-up + >= 50 timed steps
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--style", default=None, help="override the config's style (ST | TS | None)")
    ap.add_argument("--stack-c", default="64,128,256")
    ap.add_argument("--stack-t", default="32,64,128")
    ap.add_argument("--shape", default=None,
                    help="T,V,classes,style,graph for an ad-hoc training workload, e.g. 64,46,14,TS,graph.LMDHG")
    args = ap.parse_args()
    global CFG
    CFG = CONFIGS[args.config]
    if args.shape:
        parts = args.shape.split(",")
        CFG = dict(CONFIGS[2], T=int(parts[0]), V=int(parts[1]), cls=int(parts[2]), style=parts[3] if parts[3] != "None" else None,
                   graph=parts[4] if len(parts) > 4 else "graph.SHRE")
    if args.style:
        CFG = dict(CFG, style=None if args.style == "None" else args.style)
        if args.config in CONFIGS and CFG["workload"] == CONFIGS[args.config]["workload"]:
            CFG["workload"] += f" [style {args.style}]"
    if args.impl == "reference":
        return run_reference(args)
    env = Env(args)
    try:
        kind = CFG["kind"]
        if kind == "train":
            run_train(env, CFG, args)
        elif kind == "infer":
            run_infer(env, CFG, args)
        elif kind == "ensemble":
            run_ensemble(env, CFG, args)
        else:
            run_stack(env, CFG, args)
    finally:
        env.close()


if __name__ == "__main__":
    main()
