"""Time the tcgen05 GEMMs on the cfg2 shapes (M = 256*32*22 tokens) with CUDA events, and serve as the
target of `ncu --set full` captures
(PROF_ONCE=1 + `ncu --profile-from-start off`: one captured launch per shape).  Usage: python tools/prof_gemm.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import altformer_b200 as ab  # noqa: E402,F401
from altformer_b200 import ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
M = int(os.environ.get("M", 256 * 32 * 22))
dev = "cuda"
torch.manual_seed(0)


GRAPH = os.environ.get("GRAPH") == "1"       # time a CUDA-graph replay (small kernels: removes host launch gaps)
ONCE = os.environ.get("PROF_ONCE") == "1"   # ncu --profile-from-start off: capture exactly one launch per shape


def bench(name, fn, flops, bytes_):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if ONCE:
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(f"{name:44s} captured", flush=True)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:   # replay a captured graph of `iters` launches: no host launch cost between the kernels
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(iters):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
    else:
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / iters
    print(f"{name:44s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s  {bytes_ / us / 1e3:8.1f} GB/s", flush=True)


def mk(r, c, dt=torch.bfloat16, s=1.0):
    return (s * torch.randn(r, c, device=dev)).to(dt)


if os.environ.get("SHAPES") == "512":
    # temporal AltFormer stage of cfg2: 256 sequences x 32 tokens, D = 512, hidden 1024
    M5 = 256 * 32
    a512, a1024, a1536 = mk(M5, 512), mk(M5, 1024), mk(M5, 1536)
    w_qkv, w_proj, w_fc1, w_fc2 = mk(1536, 512, s=0.05), mk(512, 512, s=0.05), mk(1024, 512, s=0.05), mk(512, 1024, s=0.05)
    z1536, z512, z1024 = torch.zeros(1536, device=dev), torch.zeros(512, device=dev), torch.zeros(1024, device=dev)
    r512 = mk(M5, 512)
    F = lambda n, k: 2 * M5 * n * k
    Bt = lambda *cols: M5 * sum(cols) * 2
    bench("T qkv   1536x512 +bias", lambda: ops.gemm_tn(a512, w_qkv, 1536, bias=z1536), F(1536, 512), Bt(512, 1536))
    bench("T proj  512x512 +bias+res", lambda: ops.gemm_tn(a512, w_proj, 512, bias=z512, residual=r512), F(512, 512), Bt(512, 512, 512))
    bench("T fc1   1024x512 +gelu+preact", lambda: ops.gemm_tn(a512, w_fc1, 1024, bias=z1024, act=ops.ACT_GELU, want_preact=True),
          F(1024, 512), Bt(512, 2048))
    bench("T fc2   512x1024 +bias+res", lambda: ops.gemm_tn(a1024, w_fc2, 512, bias=z512, residual=r512), F(512, 1024), Bt(1024, 1024))
    bench("T dX    1536->512", lambda: ops.gemm_tn(a1536, w_qkv, 512, b_mn_major=True), F(1536, 512), Bt(1536, 512))
    bench("T dao   512->512", lambda: ops.gemm_tn(a512, w_proj, 512, b_mn_major=True), F(512, 512), Bt(512, 512))
    bench("T dpre  512->1024 gelu_bwd", lambda: ops.gemm_tn(a512, w_fc2, 1024, b_mn_major=True, act=ops.ACT_GELU_BWD, aux=a1024),
          F(512, 1024), Bt(512, 2048))
    bench("T dx1   1024->512", lambda: ops.gemm_tn(a1024, w_fc1, 512, b_mn_major=True), F(1024, 512), Bt(1024, 512))
    g1, g2, g3, g4 = torch.zeros(1536, 512, device=dev), torch.zeros(512, 512, device=dev), torch.zeros(1024, 512, device=dev), torch.zeros(512, 1024, device=dev)
    bench("T dW    1536x512 +dbias", lambda: ops.gemm_dw(a1536, a512, g1, dbias=z1536), F(1536, 512), Bt(1536, 512))
    bench("T dW    512x512 +dbias", lambda: ops.gemm_dw(r512, a512, g2, dbias=z512), F(512, 512), Bt(512, 512))
    bench("T dW    1024x512 +dbias", lambda: ops.gemm_dw(a1024, a512, g3, dbias=z1024), F(1024, 512), Bt(1024, 512))
    bench("T dW    512x1024 +dbias", lambda: ops.gemm_dw(r512, a1024, g4, dbias=z512), F(512, 1024), Bt(512, 1024))
    gam5, bet5 = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    bench("T layernorm fwd D=512", lambda: ops.layernorm_fwd(a512, gam5, bet5, 1e-6), 0, Bt(512, 512))
    y5, mean5, rstd5 = ops.layernorm_fwd(a512, gam5, bet5, 1e-6)
    dg5, db5 = torch.zeros(512, device=dev), torch.zeros(512, device=dev)
    bench("T layernorm bwd D=512 (+dres)", lambda: ops.layernorm_bwd(a512, a512, gam5, mean5, rstd5, dg5, db5, dres=a512), 0, Bt(512, 512, 512, 512))
    print("done512")
    sys.exit(0)

x256, x512 = mk(M, 256), mk(M, 512)
wqkv, wproj, wfc1, wfc2 = mk(768, 256, s=0.05), mk(256, 256, s=0.05), mk(512, 256, s=0.05), mk(256, 512, s=0.05)
bq, b256, b512 = torch.zeros(768, device=dev), torch.zeros(256, device=dev), torch.zeros(512, device=dev)
res = mk(M, 256)
g768, g512 = mk(M, 768), mk(M, 512)
dW = torch.zeros(768, 256, device=dev)

bench("qkv   tn 768x256 +bias", lambda: ops.gemm_tn(x256, wqkv, 768, bias=bq), 2 * M * 768 * 256, M * (256 + 768) * 2)
bench("proj  tn 256x256 +bias+residual", lambda: ops.gemm_tn(x256, wproj, 256, bias=b256, residual=res), 2 * M * 256 * 256, M * 768 * 2)
bench("fc1   tn 512x256 +bias+gelu+preact", lambda: ops.gemm_tn(x256, wfc1, 512, bias=b512, act=ops.ACT_GELU, want_preact=True),
      2 * M * 512 * 256, M * (256 + 1024) * 2)
bench("fc2   tn 256x512 +bias+residual", lambda: ops.gemm_tn(x512, wfc2, 256, bias=b256, residual=res), 2 * M * 256 * 512, M * (512 + 512) * 2)
bench("dX    mn 768->256", lambda: ops.gemm_tn(g768, wqkv, 256, b_mn_major=True), 2 * M * 768 * 256, M * (768 + 256) * 2)
bench("dpre  mn 256->512 gelu_bwd", lambda: ops.gemm_tn(x256, wfc2, 512, b_mn_major=True, act=ops.ACT_GELU_BWD, aux=x512),
      2 * M * 256 * 512, M * (256 + 1024) * 2)
bench("dx1   mn 512->256 (fc1 dX)", lambda: ops.gemm_tn(g512, wfc1, 256, b_mn_major=True), 2 * M * 512 * 256, M * (512 + 256) * 2)
bench("dao   mn 256->256 (proj dX)", lambda: ops.gemm_tn(x256, wproj, 256, b_mn_major=True), 2 * M * 256 * 256, M * (256 + 256) * 2)
x128e, wemb = mk(M, 128), mk(256, 128, s=0.05)
bench("embed tn 256x128 +bias", lambda: ops.gemm_tn(x128e, wemb, 256, bias=b256), 2 * M * 256 * 128, M * (128 + 256) * 2)
bench("dW    768x256", lambda: ops.gemm_dw(g768, x256, dW), 2 * M * 768 * 256, M * 1024 * 2)
db768, db512, db256 = torch.zeros(768, device=dev), torch.zeros(512, device=dev), torch.zeros(256, device=dev)
dW2, dW3, dW4 = torch.zeros(256, 256, device=dev), torch.zeros(512, 256, device=dev), torch.zeros(256, 512, device=dev)
bench("dW    768x256 +dbias (qkv)", lambda: ops.gemm_dw(g768, x256, dW, dbias=db768), 2 * M * 768 * 256, M * 1024 * 2)
bench("dW    256x256 +dbias (proj)", lambda: ops.gemm_dw(res, x256, dW2, dbias=db256), 2 * M * 256 * 256, M * 512 * 2)
bench("dW    512x256 +dbias (fc1)", lambda: ops.gemm_dw(g512, x256, dW3, dbias=db512), 2 * M * 512 * 256, M * 768 * 2)
bench("dW    256x512 +dbias (fc2)", lambda: ops.gemm_dw(res, x512, dW4, dbias=db256), 2 * M * 512 * 256, M * 768 * 2)
x128 = mk(M, 128)
wconv = mk(128, 9 * 128, s=0.03)
bench("conv  9x1 128->128", lambda: ops.gemm_tn(x128, wconv, 128, k_per_tap=128, taps=9, tap_row_stride=22, tap_pad=4, rows_per_batch=704,
                                                batches=M // 704, bias=torch.zeros(128, device=dev)), 2 * M * 128 * 1152, M * 256 * 2)
print("done")

# attention (cfg2 spatial stage: 8192 sequences of 22 tokens, 8 heads x 32; temporal: 256 x 32 tokens, 8 x 64)
for (B, L, H, dh) in ((256 * 32, 22, 8, 32), (256, 32, 8, 64)):
    D = H * dh
    qkv = mk(B * L, 3 * D)
    do = mk(B * L, D)
    fl = 4 * B * H * L * L * dh
    bench(f"attention fwd B={B} L={L} dh={dh}", lambda: ops.attention_fwd(qkv, B, L, H), fl, B * L * 4 * D * 2)
    bench(f"attention bwd B={B} L={L} dh={dh}", lambda: ops.attention_bwd(qkv, do, B, L, H), 2.5 * fl, B * L * 7 * D * 2)
x = mk(M, 256)
gam, bet = torch.ones(256, device=dev), torch.zeros(256, device=dev)
bench("layernorm fwd D=256", lambda: ops.layernorm_fwd(x, gam, bet, 1e-6), 0, M * 256 * 4)
y, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6)
dg, db = torch.zeros(256, device=dev), torch.zeros(256, device=dev)
bench("layernorm bwd D=256 (+dres)", lambda: ops.layernorm_bwd(x, x, gam, mean, rstd, dg, db, dres=x), 0, M * 256 * 8)
bench("colsum 768", lambda: ops.colsum(g768, torch.zeros(768, device=dev)), 0, M * 768 * 2)
print("done2")
