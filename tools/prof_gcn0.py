"""Launch the gcn0 forward (cfg2 shape: N=256, T=32, V=22 -> 128 channels, bf16 out) a few times; the
target of the `ncu --set full` capture of the HBM-bound kernel group."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altformer_b200 as ab  # noqa: E402
from oracle import altformer_oracle as O  # noqa: E402  (synthetic inputs only)

N, T, V = int(os.environ.get("N", 256)), int(os.environ.get("T", 32)), int(os.environ.get("V", 22))
x, _ = O.synthetic_batch(N, T, V, 28)
mod = ab.unit_agcn(3, 128, O.spatial_graph(V)).cuda().train()
with torch.no_grad():
    mod.bn.weight.fill_(1.0)
x = x.cuda()
with torch.no_grad():
    for _ in range(5):
        y = mod.forward_skeleton(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.float().abs().mean()))

if os.environ.get("BWD") == "1":   # forward + backward (parameter gradients), CUDA events over 10 iterations
    g = torch.randn(y.shape, device="cuda").to(y.dtype)
    for _ in range(3):
        mod.zero_grad(set_to_none=True)
        mod.forward_skeleton(x).backward(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        mod.forward_skeleton(x).backward(g)
    e1.record()
    torch.cuda.synchronize()
    print(f"gcn0 fwd+bwd (eager, incl. host launch gaps): {1e3 * e0.elapsed_time(e1) / 10:.1f} us")
