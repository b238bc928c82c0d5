"""gcn0 forward timing the way bench.py reports it: a CUDA graph of K launches over K distinct input / output buffer
sets (K x 48 MB > L2, so every launch streams its output to HBM), CUDA events around one replay -> us per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altformer_b200 as ab  # noqa: E402


def time_gcn0(mod, xs, reps=5):
    dev = xs[0].device
    with torch.no_grad():
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for x in xs[:2]:
                mod.forward_skeleton(x)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            keep = [mod.forward_skeleton(x) for x in xs]
        times = []
        for _ in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            times.append(1e3 * e0.elapsed_time(e1) / len(xs))
    del keep
    return sorted(times[2:])[len(times[2:]) // 2]


if __name__ == "__main__":
    from oracle import altformer_oracle as O
    N, T, V = int(os.environ.get("N", 256)), int(os.environ.get("T", 32)), int(os.environ.get("V", 22))
    K = int(os.environ.get("K", 8))
    mod = ab.unit_agcn(3, 128, O.spatial_graph(V)).cuda().train()
    with torch.no_grad():
        mod.bn.weight.fill_(1.0)
    xs = [O.synthetic_batch(N, T, V, 28, 100 + i)[0].cuda() for i in range(K)]
    for train in (True, False):
        mod.train(train)
        us = time_gcn0(mod, xs)
        nbytes = N * T * V * (12 + 256)
        print(f"gcn0 fwd N={N} T={T} V={V} train={train}: {us:.2f} us per launch = {nbytes / us / 1e3:.0f} GB/s "
              f"(fused={os.environ.get('AFB_GCN0_FUSED', '1')})")
    if not os.environ.get("AFB_GCN0_STAMPS"):
        sys.exit(0)   # the stamps exist only in a library built with AFB_GCN0_STAMPS=1 (python st-gcn-altformer_b200/build.py --force)
    # phase stamps of one plain launch (globaltimer, thread 0 of each CTA), relative to the earliest CTA entry
    import ctypes as C
    import numpy as np
    mod.train(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    names = ["entry", "loads staged", "coefficients", "operands (S1)", "M done (S2)", "r rows (S4)", "moments mma", "moments posted",
             "barrier passed", "slot sums", "weights folded", "stores issued", "stores drained"]
    for rep in range(2):
        flush.zero_()
        with torch.no_grad():
            mod.forward_skeleton(xs[rep])
        torch.cuda.synchronize()
        ctas = 296
        buf = (C.c_uint64 * (16 * ctas))()
        rc = ab._lib.lib().afb_gcn0_fused_stamps(buf, ctas)
        st = np.frombuffer(buf, dtype=np.uint64).reshape(ctas, 16).astype(np.int64)
        busy = st[:N] if N < ctas else st
        t0 = st[:, 0].min()
        print(f"-- phase stamps, launch {rep} (ns after the first CTA's entry; CTAs with a sample: min / median / max)")
        for k, nm in enumerate(names):
            col = (busy[:, k] - t0) if k in (3, 4, 5, 6, 11, 12) else (st[:, k] - t0)
            print(f"   {k:2d} {nm:16s} {col.min():7d} {int(np.median(col)):7d} {col.max():7d}")
