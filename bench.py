"""Benchmark of the north-star hot path: SHREC'17-shape ST_GCN_AltFormer training step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = zero_grad -> forward -> cross-entropy -> backward -> gradient all-reduce (N > 1) -> AdamW on one
batch of synthetic skeletons (BASELINE.json configs[1]: batch 256 per GPU, T=32, V=22, 28 classes, bf16
compute with fp32 master weights / accumulation, style 'ST').  Weak scaling: every rank processes its own
256-sample shard; gradients are averaged with one NCCL all-reduce.

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the
public API with pinned-host inputs (H2D inside the timed region) and a D2H read of the loss every step.
`roofline` is the gcn0 (unit_agcn 3->128) forward, the HBM-bound kernel group the metric names;
`roofline_tensor` is the whole step against the measured dense bf16 peak.  `--impl reference` times the CPU
oracle port of the reference step (the reference itself is pure Python and absent on the GPU box).
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

CFG = dict(T=32, V=22, cls=28, per_gpu_batch=256, style="ST")
METRIC = "AltFormer train seq/s (SHREC shape)"


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


def flops_per_sample_fwd(T, V, cls):
    """transformer block 16 D^2 M + 4 L D M, embeds, tcn0, gcn0, head (style ST)."""
    M1, M2 = T * V, T
    blk = lambda D, M, L: 16 * D * D * M + 4 * L * D * M  # noqa: E731
    f = 6 * blk(256, M1, V) + 6 * blk(512, M2, T)
    f += 2 * 128 * 256 * M1 + 2 * 256 * 512 * M2 + 2 * 512 * cls
    f += 2 * 128 * 128 * 9 * M1 + 7.13e6 * (T * V) / (32 * 22)
    return f


def synthetic_batch(N, T, V, cls, seed=1234):
    """BASELINE.md 4 / SURVEY 8d inputs: 0.2*randn skeletons, palm-centred (Hand_Dataset.py:61), uniform labels.
    (Own copy: the measured arm does not import anything from oracle/.)"""
    g = torch.Generator().manual_seed(seed)
    x = 0.2 * torch.randn(N, T, V, 3, generator=g)
    x = x - x[:, :1, 1:2, :]
    y = torch.randint(0, cls, (N,), generator=g)
    return x, y


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline is meant to use every core it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return n


def run_reference(args):
    """CPU oracle port of the reference train step (train_sttran.py:89-102,185-191), all host threads."""
    from oracle import altformer_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, V, cls = CFG["T"], CFG["V"], CFG["cls"]
    B = 32  # bounded sample of the 256-sample step: seq/s on CPU is batch-insensitive
    use_all_host_threads()
    torch.manual_seed(0)
    model = O.OracleModel(O.random_state(O.model_spec(3, cls, T, V), 0), O.spatial_graph(V), CFG["style"]).train()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.1)
    x, y = O.synthetic_batch(B, T, V, cls)
    crit = torch.nn.CrossEntropyLoss()

    def step():
        loss = crit(model(x), y)
        model.zero_grad()
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(max(args.warmup, 1)):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = B / (ms / 1e3)
    cores = torch.get_num_threads()
    sample = f"batch {B} of the 256-sample step (fwd+CE+bwd+AdamW), fp32, {cores} threads"
    emit(({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: SHREC-shape ST_GCN_AltFormer(style ST) train step, T=32 V=22 28 classes", "cpu_batch": B},
        "cpu_baseline": {"value": val, "unit": "seq/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def timed(fn, steps, dist_on, dev):
    """barrier + sync, K steps between CUDA events, sync + barrier; returns max-over-ranks ms/step."""
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    if dist_on:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t)
    return ms


def gcn0_roofline(model, x, dev, pk, iters=20):
    """gcn0 forward alone (scores + finalize + apply launches), L2 flushed between launches."""
    import altformer_b200 as ab  # noqa: F401
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
    N, T, V, _ = x.shape
    times = []
    with torch.no_grad():
        # the two gcn0 launches are replayed from a CUDA graph so the events see GPU time, not Python launch gaps
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model.gcn0.forward_skeleton(x)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            model.gcn0.forward_skeleton(x)
        for i in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            if i >= 3:
                times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    alg_bytes = N * T * V * (3 * 4 + 128 * 2)
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "gcn0 = unit_agcn(3->128) forward (gcn0_scores [+finalize in its last CTA] + gcn0_apply)",
            "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"], "frac_of_8TBps_nominal": achieved / 8000.0,
            "peak_source": pk["src"],
            # dram__bytes_read.sum + dram__bytes_write.sum of the two kernels, ncu --set full at this batch
            # (profiles/r01_ncu_gcn0_v3.txt): 6.5 MB read; the 46 MB output lands in L2 and is written back while the
            # next kernel runs (45.1 MB of write-back observed there).  Only meaningful for the default batch of 256.
            "traffic": 51.6e6 if N == 256 and T == 32 and V == 22 else None,
            "alg_bytes_per_launch": alg_bytes, "ms_per_launch": ms}


def step_kernel_rooflines(M, dev, pk, iters=10):
    """The kernels that carry the step's time (spatial stage, D = 256, M = per-GPU tokens), each timed live with CUDA
    events over `iters` back-to-back launches on tensors far larger than L2: algorithmic bytes / time vs the HBM peak."""
    from altformer_b200 import ops
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
    return out


def cpu_baseline(budget_s=20.0):
    from oracle import altformer_oracle as O
    use_all_host_threads()
    T, V, cls, B = CFG["T"], CFG["V"], CFG["cls"], 32
    model = O.OracleModel(O.random_state(O.model_spec(3, cls, T, V), 0), O.spatial_graph(V), CFG["style"]).train()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.1)
    x, y = O.synthetic_batch(B, T, V, cls)
    crit = torch.nn.CrossEntropyLoss()
    times, t_start = [], time.perf_counter()
    while len(times) < 2 or (time.perf_counter() - t_start < budget_s and len(times) < 12):
        t0 = time.perf_counter()
        loss = crit(model(x), y)
        model.zero_grad()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * statistics.median(times[1:])
    cores = torch.get_num_threads()
    return {"value": B / (ms / 1e3), "unit": "seq/s", "cores": cores, "kind": "port",
            "sample": f"{len(times) - 1} train steps of batch {B} (of the 256-sample step), oracle port, fp32, {cores} threads"}


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


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=CFG["per_gpu_batch"])
    ap.add_argument("--shape", default=None,
                    help="T,V,classes,style,graph for an ad-hoc workload, e.g. 64,46,14,TS,graph.LMDHG (BASELINE configs[3]); "
                         "the default (and the line the driver records) is configs[1]")
    args = ap.parse_args()
    graph_name = "graph.SHRE"
    if args.shape:
        parts = args.shape.split(",")
        CFG.update(T=int(parts[0]), V=int(parts[1]), cls=int(parts[2]), style=parts[3] if parts[3] != "None" else None)
        graph_name = parts[4] if len(parts) > 4 else graph_name
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import altformer_b200 as ab
    from altformer_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist_on = world > 1
    if dist_on:
        dist.init_process_group("nccl", device_id=dev)
    T, V, cls, B = CFG["T"], CFG["V"], CFG["cls"], args.batch
    pk = peaks()

    torch.manual_seed(0)
    model = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style=CFG["style"], graph=graph_name,
                                graph_args={"labeling_mode": "spatial"}).to(dev)
    with torch.no_grad():  # the reference's init makes gcn0 invisible (bn gamma 1e-6); use a live one for a fair workload
        model.gcn0.bn.weight.fill_(1.0)
    trainer = ab.DataParallelTrainer(model, use_graph=not args.no_graph)
    x_cpu, y_cpu = synthetic_batch(B, T, V, cls, 1234 + rank)
    x_pin, y_pin = x_cpu.pin_memory(), y_cpu.pin_memory()
    x_dev, y_dev = x_pin.to(dev), y_pin.to(dev)

    def step_resident():
        trainer.step(x_dev, y_dev)

    # e2e: every step uploads its batch from pinned host memory (straight into the step's input buffers) and
    # reads its loss back (as get_acc does, train_sttran.py:105-109).  The read-back is asynchronous and consumed
    # one step later, so the host enqueues step i+1 while step i runs; the timed region ends with a full sync.
    host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "last": float("nan")}

    def step_e2e():
        k = e2e_state["i"] & 1
        loss, _ = trainer.step(x_pin, y_pin)
        host_loss[k].copy_(loss, non_blocking=True)
        loss_ready[k].record()
        if e2e_state["i"] > 0:
            loss_ready[k ^ 1].synchronize()
            e2e_state["last"] = float(host_loss[k ^ 1])
        e2e_state["i"] += 1

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # count our kernel launches in one eager-equivalent step
    ops.LAUNCHES[0] = 0
    if args.no_graph:
        step_resident()
        launches = ops.LAUNCHES[0]
    else:
        launches = trainer.launches_per_step
    with ClockSampler(local) as clk:
        ms = timed(step_resident, args.steps, dist_on, dev)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps, dist_on, dev)

    if rank == 0:
        gb = B * world
        roof = gcn0_roofline(model, x_dev, dev, pk)
        flops_step = 3.0 * flops_per_sample_fwd(T, V, cls) * B
        tf = flops_step / (ms * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": gb / (ms * 1e-3), "unit": "seq/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("configs[1]: SHREC'17-shape ST_GCN_AltFormer(style ST) fwd+bwd+AdamW training step, T=32 V=22 28 classes"
                                    if not args.shape else f"ad-hoc --shape {args.shape}: ST_GCN_AltFormer fwd+bwd+AdamW training step"),
                       "per_gpu_batch": B, "global_batch": gb, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                       "l2": "no explicit flush: the step streams >5 GB of activations per GPU, far beyond the 126 MB L2"},
            "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "seq/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_pin.numel() * 4 + y_pin.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "roofline": roof,
            # the kernels the step actually spends its time in (gcn0 above is the metric's kernel but <1 % of the step)
            "roofline_step_kernels": (step_kernel_rooflines(B * T * V, dev, pk) if V == 22 and not args.shape else None),
            "roofline_tensor": {"bound": "tensor", "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"],
                                "peak_source": pk["src"] + " sustained", "note": "whole step, 3 x forward FLOPs (SURVEY 8d), per GPU"},
        }
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        emit(out)
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
