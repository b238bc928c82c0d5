"""Attention forward: tcgen05 / TMEM / TMA kernel (attention_tc.cu, AFB_ATTN_TC=1 | 2) vs the warp-level MMA kernel
(attention_mma.cu) vs a torch fp32 softmax reference, on the shapes the models use.  Prints rel-L2 errors and CUDA-event
timings (us per launch, 10 launches back to back on buffers > L2 at the large shapes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from altformer_b200 import ops  # noqa: E402


def ref(qkv, B, L, H):
    D = qkv.shape[1] // 3
    dh = D // H
    q, k, v = qkv.float().view(B, L, 3, H, dh).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * L, D)


def run(mode, qkv, B, L, H, out_scale=None):
    if mode == "mma":
        os.environ["AFB_ATTN_TC"] = "0"
    else:
        os.environ["AFB_ATTN_TC"] = "1" if mode == "tc" else "2"
    return ops.attention_fwd(qkv, B, L, H, out_scale)


def time_us(mode, qkv, B, L, H, iters=10):
    for _ in range(3):
        run(mode, qkv, B, L, H)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        run(mode, qkv, B, L, H)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    dev = "cuda"
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())  # noqa: E731
    bad = 0
    # (B, L, heads, dh)
    for B, L, H, dh in [(7, 22, 8, 32), (64, 22, 8, 32), (33, 32, 8, 32), (5, 9, 4, 32), (6, 46, 8, 32), (9, 64, 8, 32),
                        (10, 32, 8, 64), (3, 46, 8, 64), (4, 64, 4, 64), (8192, 22, 8, 32), (5632, 32, 8, 64), (2944, 64, 8, 32)]:
        D = H * dh
        qkv = (torch.randn(B * L, 3 * D, device=dev) * 1.5).to(torch.bfloat16)
        want = ref(qkv, B, L, H)
        keep = (torch.rand(B, device=dev) > 0.3).float() / 0.7
        line = f"B={B:5d} L={L:2d} H={H} dh={dh}:"
        for mode in ("mma", "tc", "tcs"):   # tc: P through tensor memory; tcs: P through shared memory
            try:
                got = run(mode, qkv, B, L, H)
                torch.cuda.synchronize()
                e = rel(got, want)
                got2 = run(mode, qkv, B, L, H, keep)
                e2 = rel(got2, want * keep.repeat_interleave(L)[:, None])
                ok = e < 1e-2 and e2 < 1e-2
                bad += 0 if ok or mode == "mma" else 1
                line += f"  {mode} {e:.2e}/{e2:.2e}{'' if ok else ' FAIL'}"
                if B >= 1024:
                    line += f" {time_us(mode, qkv, B, L, H):6.1f}us"
            except Exception as ex:  # noqa: BLE001
                bad += 1
                line += f"  {mode} ERROR {str(ex)[:80]}"
        if B >= 1024:
            line += f"   (HBM floor {B * L * 4 * D * 2 / 6542.1e3:.1f}us)"
        print(line, flush=True)
        if dh == 32 and H % 2 == 0:   # backward: tcgen05 kernel vs warp-level kernel vs torch autograd
            qr = qkv.float().requires_grad_(True)
            do = (torch.randn(B * L, D, device=dev)).to(torch.bfloat16)
            ref(qr, B, L, H).backward(do.float())
            line = f"      backward:"
            for mode in ("mma", "tc"):
                os.environ["AFB_ATTN_TC_BWD"] = "1" if mode == "tc" else "0"
                try:
                    got = ops.attention_bwd(qkv, do, B, L, H)
                    torch.cuda.synchronize()
                    e = rel(got, qr.grad)
                    parts = [rel(got[:, k * D:(k + 1) * D], qr.grad[:, k * D:(k + 1) * D]) for k in range(3)]
                    ok = e < 1e-2
                    bad += 0 if ok or mode == "mma" else 1
                    line += f"  {mode} {e:.2e} (dq {parts[0]:.1e} dk {parts[1]:.1e} dv {parts[2]:.1e}){'' if ok else ' FAIL'}"
                    if B >= 1024:
                        for _ in range(3):
                            ops.attention_bwd(qkv, do, B, L, H)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        e0.record()
                        for _ in range(10):
                            ops.attention_bwd(qkv, do, B, L, H)
                        e1.record()
                        torch.cuda.synchronize()
                        line += f" {100 * e0.elapsed_time(e1):6.1f}us"
                except Exception as ex:  # noqa: BLE001
                    bad += 1
                    line += f"  {mode} ERROR {str(ex)[:120]}"
            os.environ.pop("AFB_ATTN_TC_BWD", None)
            if B >= 1024:
                line += f"   (HBM floor {B * L * 7 * D * 2 / 6542.1e3:.1f}us)"
            print(line, flush=True)
    print("attn_tc_check:", "ok" if bad == 0 else f"{bad} FAILED")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
