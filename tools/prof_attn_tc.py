"""One shape of the attention forward / backward for ncu: MODE=tc|tcs|mma, BWD=0|1, B, L, H, DH from the environment."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from altformer_b200 import ops  # noqa: E402

B, L, H, DH = (int(os.environ.get(k, d)) for k, d in (("B", 8192), ("L", 22), ("H", 8), ("DH", 32)))
mode = os.environ.get("MODE", "tc")
bwd = os.environ.get("BWD", "0") == "1"
os.environ["AFB_ATTN_TC"] = {"mma": "0", "tc": "1", "tcs": "2"}[mode]
os.environ["AFB_ATTN_TC_BWD"] = "0" if mode == "mma" else "1"
qkv = (torch.randn(B * L, 3 * H * DH, device="cuda") * 1.5).to(torch.bfloat16)
do = torch.randn(B * L, H * DH, device="cuda").to(torch.bfloat16)
for _ in range(4):
    o = ops.attention_bwd(qkv, do, B, L, H) if bwd else ops.attention_fwd(qkv, B, L, H)
torch.cuda.synchronize()
print("ok", tuple(o.shape), float(o.float().abs().mean()))
