"""BASELINE configs[4] (cfg5): large-batch SHREC-shape inference, joint + bone + motion streams, ST and TS models per
stream, combined 0.8*ST + 0.2*TS and summed over streams (emsemble.py:215-225).  Not the bench.py line (that is
configs[1]); this reports input sequences/s and model forwards/s for DESIGN.md.

    python tools/bench_cfg5.py [batch] [iters]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altformer_b200 as ab  # noqa: E402
from oracle import altformer_oracle as O  # noqa: E402  (synthetic inputs / seeded weights only)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T, V, cls = 32, 22, 28
dev = torch.device("cuda", 0)


def build(style, seed):
    m = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style=style, graph="graph.SHRE", graph_args={"labeling_mode": "spatial"})
    m.load_state_dict(O.random_state(O.model_spec(3, cls, T, V), seed))
    return m.to(dev).eval()


models = {s: (build("ST", 10 + i), build("TS", 20 + i)) for i, s in enumerate(("joint", "bone", "motion"))}
x, _ = O.synthetic_batch(N, T, V, cls, 7)
x = x.to(dev)
for _ in range(2):
    y = ab.streams.ensemble_forward(x, models)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = ab.streams.ensemble_forward(x, models)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
assert torch.isfinite(y).all() and y.shape == (N, cls)
# small-batch cross-check of the same pipeline against per-model calls
xs = x[:64]
ref = sum(ab.streams.combine(a(s), b(s)) for (a, b), s in ((models["joint"], xs), (models["bone"], ab.streams.bone(xs)), (models["motion"], ab.streams.motion(xs))))
err = float((ab.streams.ensemble_forward(xs, models).float() - ref.float()).norm() / ref.float().norm())
print(f"cfg5: batch {N}, 3 streams x (ST, TS): {ms:.1f} ms per ensemble pass -> {N / ms * 1e3:.0f} input seq/s, "
      f"{6 * N / ms * 1e3:.0f} model forwards/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, self-check rel {err:.1e}")
