"""One TCN_GCN_unit(C, C) forward + backward at the configs[2] size (for an ncu launch list / --set full captures).
    C=128 T=32 N=1024 EXACT=1 python tools/prof_stack.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altformer_b200 as ab  # noqa: E402
from altformer_b200 import functional as AF  # noqa: E402

C, T, N, V = (int(os.environ.get(k, d)) for k, d in (("C", 128), ("T", 32), ("N", 1024), ("V", 22)))
AF.set_exact_bn_mask(os.environ.get("EXACT", "1") != "0")
dev = torch.device("cuda", 0)
A = torch.as_tensor(ab.import_class("graph.SHRE" if V == 22 else "graph.LMDHG")(labeling_mode="spatial").A, dtype=torch.float32)
torch.manual_seed(0)
unit = ab.TCN_GCN_unit(C, C, A, dropout=0.0).to(dev).train()
x = torch.randn(N * T * V, C, device=dev).to(torch.bfloat16).requires_grad_(True)


def step():
    h = unit.forward_tokens(x, (N, T, V))
    h.backward(h.detach())


for _ in range(int(os.environ.get("WARM", 2))):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print(f"TCN_GCN_unit C={C} T={T} N={N} V={V} exact={AF.exact_bn_mask()}: fwd+bwd {e0.elapsed_time(e1):.2f} ms")
