"""bf16 noise floor of the REFERENCE MATH itself: the oracle (pinned to the reference, tests/golden) run under
torch.autocast(bfloat16) against its own fp32 run, same weights and inputs, per tensor (relative L2 of logits and of every
parameter gradient).  The whole-model bf16 parity cases compare the CUDA path's error tensor by tensor with this floor
(tests/golden/autocast_floor.json): a deep pre-LN transformer amplifies the 2^-9 operand rounding of ANY bf16 pipeline --
PyTorch's own included -- to 1e-2 .. 2e-1 on individual early-layer gradients, so "<= 1e-2 against fp32" is a per-module
bar (met, tools/gpu_diag_modules.py) and "no worse than the reference under autocast" is the whole-model bar.

    python tools/autocast_floor.py [--full]   # (re)generates entries of tests/golden/autocast_floor.json (CPU, ~2 min;
                                              #  --full adds the batch-256 configs[1] case, ~15 min)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import altformer_oracle as O  # noqa: E402

# (key, style, N, T, V, classes, weight seed, batch seed, loss): "cot" = sum(y * fixed random cotangent) as in
# tools/gpu_diag_modules.py:grp_model; "ce" = the training loss of tests/test_gpu_parity.py::test_cfg2_train_step_against_oracle
CASES = [("ST_N4_T32_V22", "ST", 4, 32, 22, 28, 61, 161, "cot"), ("TS_N4_T32_V22", "TS", 4, 32, 22, 28, 61, 161, "cot"),
         ("TS_N2_T64_V46", "TS", 2, 64, 46, 14, 61, 161, "cot")]
FULL = [("cfg2_ST_N256", "ST", 256, 32, 22, 28, 5, 123, "ce")]     # --full: BASELINE configs[1] size (about 15 min of CPU)


def run(style, N, T, V, cls, wseed, bseed, loss, autocast):
    A = O.spatial_graph(V)
    st = O.random_state(O.model_spec(3, cls, T, V), wseed)
    x, labels = O.synthetic_batch(N, T, V, cls, bseed)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in st.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        y = O.model_forward(x, params, A, style, True)
    if loss == "ce":
        torch.nn.functional.cross_entropy(y.float(), labels).backward()
    else:
        cot = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
        (y.float() * cot).sum().backward()
    return y.detach().float(), params


def main():
    path = os.path.join(ROOT, "tests", "golden", "autocast_floor.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for key, *case in (FULL if "--full" in sys.argv else CASES):
        y0, p0 = run(*case, False)
        y1, p1 = run(*case, True)
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))  # noqa: E731
        grads = {k: rel(p1[k].grad, p0[k].grad) for k in p0 if getattr(p0[k], "grad", None) is not None and float(p0[k].grad.norm()) > 0}
        out[key] = {"logits": rel(y1, y0), "grads": grads}
        es = sorted(grads.values())
        print(f"{key}: logits {out[key]['logits']:.3e}; grads median {es[len(es) // 2]:.3e} worst {es[-1]:.3e}", flush=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
