"""On-GPU diagnostics: every check compares one C-ABI kernel with a plain torch reference and prints
PASS/FAIL plus error metrics.  Each group runs in its own process (tools/gpu_run_all.sh) so that a
faulting kernel cannot poison the others.  Usage: python tools/gpu_diag.py <group> [...]"""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import altformer_b200 as ab  # noqa: E402
from altformer_b200 import functional as AF  # noqa: E402
from altformer_b200 import ops  # noqa: E402

DEV = "cuda"
RESULTS = []


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)), float((a - b).norm() / b.norm().clamp_min(1e-30))


def report(name, got, ref, tol):
    torch.cuda.synchronize()
    bad = not torch.isfinite(got.float()).all()
    e_inf, e_l2 = rel(got, ref)
    ok = (not bad) and e_inf <= tol and e_l2 <= tol
    RESULTS.append((name, ok))
    print(f"{'PASS' if ok else 'FAIL'} {name:58s} rel_inf={e_inf:.3e} rel_l2={e_l2:.3e} tol={tol:g}{' NONFINITE' if bad else ''}", flush=True)
    return ok


def check(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        RESULTS.append((fn.__name__, False))
        print(f"FAIL {fn.__name__}: EXCEPTION {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def g(*shape, seed=0, scale=1.0, dtype=torch.float32):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    return (scale * torch.randn(*shape, generator=gen)).to(DEV).to(dtype)


# --------------------------------------------------------------------------------------------
def grp_simt():
    def simt_basic():
        a, b = g(70, 50, seed=1), g(33, 50, seed=2)
        bias = g(33, seed=3)
        c = ops.gemm_simt(a, b, 70, 33, 50, (50, 1), (50, 1), bias=bias)
        report("gemm_simt f32 70x33x50 +bias", c, a @ b.t() + bias, 1e-5)
        c = ops.gemm_simt(a, b, 50, 50, 70, (1, 50), (1, 50)) if False else None
        at = g(50, 70, seed=4)
        c = ops.gemm_simt(at, b, 70, 33, 50, (1, 70), (50, 1))
        report("gemm_simt A transposed", c, at.t() @ b.t(), 1e-5)
        ab16, bb16 = a.bfloat16(), b.bfloat16()
        c = ops.gemm_simt(ab16, bb16, 70, 33, 50, (50, 1), (50, 1), out_dtype=torch.bfloat16)
        report("gemm_simt bf16", c, ab16.float() @ bb16.float().t(), 1e-2)
    check(simt_basic)


# --------------------------------------------------------------------------------------------
def grp_elementwise():
    def cast_copy():
        x = g(1000, 37, seed=1)
        report("cast f32->bf16", ops.cast(x, torch.bfloat16), x.bfloat16(), 1e-6)
        report("cast bf16->f32", ops.cast(x.bfloat16(), torch.float32), x.bfloat16().float(), 0)
        w = g(48, 20, seed=2)
        report("cast_transpose", ops.cast_transpose(w), w.t().bfloat16(), 1e-6)
        dst = torch.zeros(10, 64, device=DEV)
        ops.copy2d(w, dst, 10, 20, 20, 64, src_off=40, dst_off=3)
        ref = torch.zeros(10, 64, device=DEV)
        ref[:, 3:23] = w[2:12]
        report("copy2d strided", dst, ref, 0)
        x3 = g(9, 16, seed=3)
        hi = x3.bfloat16()
        lo = (x3 - hi.float()).bfloat16()
        report("split3 A", ops.split3(x3, 0), torch.cat([hi, lo, hi], 1), 0)
        report("split3 B", ops.split3(x3, 1), torch.cat([hi, hi, lo], 1), 0)
        report("split3 rows", ops.split3(x3, 2).view(27, 16), torch.cat([hi, hi, lo], 0), 0)
        wc = g(8, 4, 9, 1, seed=4)
        f, b = ops.conv_weight_pack(wc)
        report("conv_pack fwd", f, wc[..., 0].permute(0, 2, 1).reshape(8, 36).bfloat16(), 0)
        report("conv_pack bwd", b, wc[..., 0].flip(2).permute(1, 2, 0).reshape(4, 72).bfloat16(), 0)
    check(cast_copy)

    def layernorm():
        for D in (128, 256, 512):
            for dtp, tol in ((torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)):
                x = g(333, D, seed=D, dtype=dtp)
                gam, bet = g(D, seed=1) * 0.3 + 1, g(D, seed=2) * 0.1
                y, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6)
                xr = x.float().requires_grad_(True)
                yr = torch.nn.functional.layer_norm(xr, (D,), gam, bet, 1e-6)
                report(f"layernorm_fwd D={D} {dtp}", y, yr.detach(), tol)
                dy = g(333, D, seed=7, dtype=dtp)
                dres = g(333, D, seed=8, dtype=dtp)
                dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
                dx = ops.layernorm_bwd(dy, x, gam, mean, rstd, dg, db, dres=dres)
                gam_r = gam.clone().requires_grad_(True)
                bet_r = bet.clone().requires_grad_(True)
                yr = torch.nn.functional.layer_norm(xr, (D,), gam_r, bet_r, 1e-6)
                yr.backward(dy.float())
                report(f"layernorm_bwd dx D={D} {dtp}", dx, xr.grad + dres.float(), tol)
                report(f"layernorm_bwd dgamma D={D} {dtp}", dg, gam_r.grad, tol)
                report(f"layernorm_bwd dbeta D={D} {dtp}", db, bet_r.grad, tol)
                xr.grad = None
    check(layernorm)

    def batchnorm():
        N, T, V, Cc = 3, 8, 22, 128
        M = N * T * V
        for dtp, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
            x = g(M, Cc, seed=1, dtype=dtp)
            gam, bet = g(Cc, seed=2) * 0.3 + 1, g(Cc, seed=3) * 0.1
            rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
            acc = ops.colstats(x)
            report(f"colstats sum {dtp}", acc[0], x.double().sum(0), 1e-6)
            report(f"colstats sumsq {dtp}", acc[1], (x.double() ** 2).sum(0), 1e-6)
            st = ops.bn_finalize(acc, M, gam, bet, rm, rv, 0.1, 1e-5, True)
            res = g(M, Cc, seed=4, dtype=dtp)
            y, y2 = ops.bn_act_fwd(x, st[2], st[3], True, res_pre=res, T=T, V=V, want_perm=True)
            xr = x.float().requires_grad_(True)
            rr = res.float().requires_grad_(True)
            gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
            rm2, rv2 = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
            x4 = xr.view(N, T, V, Cc).permute(0, 3, 1, 2)
            yr = torch.relu(torch.nn.functional.batch_norm(x4, rm2, rv2, gr, br, True, 0.1, 1e-5) + rr.view(N, T, V, Cc).permute(0, 3, 1, 2))
            yr_tok = yr.permute(0, 2, 3, 1).reshape(M, Cc)
            report(f"bn_act_fwd {dtp}", y, yr_tok.detach(), tol)
            report(f"bn_act_fwd perm {dtp}", y2, yr.permute(0, 3, 2, 1).reshape(M, Cc).detach(), tol)
            report(f"bn running_mean {dtp}", rm, rm2, 1e-4)
            report(f"bn running_var {dtp}", rv, rv2, 1e-4)
            dy = g(M, Cc, seed=5, dtype=dtp)
            dy2 = g(M, Cc, seed=6, dtype=dtp)
            dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
            dx, dres = ops.bn_bwd(dy, dy2, x, st, gam, bet, True, True, dg, db, res_pre=res, want_dres=True, T=T, V=V)
            tot = dy.float() + dy2.float().view(N, V, T, Cc).permute(0, 2, 1, 3).reshape(M, Cc)
            yr_tok.backward(tot)
            report(f"bn_bwd dx {dtp}", dx, xr.grad, tol)
            report(f"bn_bwd dres {dtp}", dres, rr.grad, tol)
            report(f"bn_bwd dgamma {dtp}", dg, gr.grad, tol)
            report(f"bn_bwd dbeta {dtp}", db, br.grad, tol)
        x = g(500, 28 * 4, seed=9)
        out = torch.zeros(112, device=DEV)
        rs = torch.rand(5, device=DEV)
        ops.colsum(x, out, rs, 100)
        report("colsum row_scale C=112", out, (x * rs.repeat_interleave(100)[:, None]).sum(0), 1e-5)
        x = g(64, 22 * 256, seed=10, dtype=torch.bfloat16)
        out = torch.zeros(22 * 256, device=DEV)
        ops.colsum(x, out)
        report("colsum C=5632 bf16", out, x.float().sum(0), 1e-5)
    check(batchnorm)

    def pools_ce_adam():
        for dtp, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
            x = g(6 * 22, 256, seed=1, dtype=dtp)
            y = ops.pool_mean_fwd(x, 6, 22)
            report(f"pool_mean_fwd {dtp}", y, x.float().view(6, 22, 256).mean(1), tol)
            dy = g(6, 256, seed=2, dtype=dtp)
            report(f"pool_mean_bwd {dtp}", ops.pool_mean_bwd(dy, 6, 22), (dy.float() / 22)[:, None, :].expand(6, 22, 256).reshape(-1, 256), tol)
            y, arg = ops.pool_max_fwd(x, 6, 22)
            ref, idx = x.float().view(6, 22, 256).max(1)
            report(f"pool_max_fwd {dtp}", y, ref, tol)
            dref = torch.zeros(6, 22, 256, device=DEV).scatter_(1, arg.long()[:, None, :], dy.float()[:, None, :])
            report(f"pool_max_bwd {dtp}", ops.pool_max_bwd(dy, arg, 6, 22), dref.view(-1, 256), tol)
        for cls in (14, 28):
            z = g(37, cls, seed=3).requires_grad_(True)
            lab = torch.randint(0, cls, (37,), device=DEV)
            loss, dl = ops.softmax_ce(z.detach(), lab)
            lr_ = torch.nn.functional.cross_entropy(z, lab)
            lr_.backward()
            report(f"softmax_ce loss C={cls}", loss, lr_.detach(), 1e-5)
            report(f"softmax_ce dlogits C={cls}", dl, z.grad, 1e-5)
            report(f"scale_rows any C={cls}", ops.scale_rows(dl, torch.full((1,), 2.0, device=DEV), 37), 2 * dl, 1e-6)
        n = 4096 + 8
        p = g(n, seed=4)
        grad = g(n, seed=5)
        m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        pr = p.clone().requires_grad_(True)
        opt = torch.optim.AdamW([pr], lr=2e-4, weight_decay=0.1)
        step = torch.ones((), device=DEV, dtype=torch.int32)
        pl = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
        for it in range(3):
            gg = grad * (it + 1)
            ops.adamw(p, gg, m, v, pl, step, 2e-4, 0.9, 0.999, 1e-8, 0.1)
            pr.grad = gg.clone()
            opt.step()
        report("adamw 3 steps", p, pr.detach(), 1e-6)
        report("adamw bf16 shadow", pl, p.bfloat16(), 0)
        report("adamw step counter", step.float(), torch.tensor(4.0), 0)
    check(pools_ce_adam)

    def streams():
        x = g(3, 5, 22, 3, seed=1)
        from oracle import altformer_oracle as O
        report("bone stream", ab.streams.bone(x), O.bone_stream(x.cpu()).to(DEV), 1e-6)
        report("motion stream", ab.streams.motion(x), O.motion_stream(x.cpu()).to(DEV), 1e-6)
        a, b = g(7, 14, seed=2), g(7, 14, seed=3)
        report("combine", ab.streams.combine(a, b), 0.8 * a + 0.2 * b, 1e-6)
    check(streams)


# --------------------------------------------------------------------------------------------
def _mm_ref(a, b):
    return a.float() @ b.float().t()


def grp_gemm_tn():
    def probe_identity():
        # B = first 64 columns selector: C[m, n] = A[m, n]: exposes swizzle / descriptor mistakes column by column
        a = g(128, 64, seed=1, dtype=torch.bfloat16)
        b = torch.eye(64, device=DEV, dtype=torch.bfloat16)
        c = ops.gemm_tn(a, b, 64, out_dtype=torch.float32)
        torch.cuda.synchronize()
        ok = report("gemm_tn probe identity 128x64x64", c, a.float(), 1e-6)
        if not ok:
            good_cols = ((c - a.float()).abs().max(0).values < 1e-6).nonzero().flatten().tolist()
            good_rows = ((c - a.float()).abs().max(1).values < 1e-6).nonzero().flatten().tolist()
            print("   correct columns:", good_cols)
            print("   correct rows   :", good_rows[:40], "...")
            print("   c[0,:8]", c[0, :8].tolist(), " a[0,:8]", a[0, :8].float().tolist())
    check(probe_identity)

    def basic_shapes():
        for (M, N, K) in ((128, 64, 64), (128, 128, 128), (256, 256, 256), (300, 256, 128), (1000, 768, 256), (4096, 512, 1024),
                          (77, 128, 1152), (20000, 1536, 512)):
            a, b = g(M, K, seed=M, dtype=torch.bfloat16), g(N, K, seed=N + 1, scale=0.1, dtype=torch.bfloat16)
            c = ops.gemm_tn(a, b, N, out_dtype=torch.float32)
            report(f"gemm_tn {M}x{N}x{K} f32 out", c, _mm_ref(a, b), 2e-3)
    check(basic_shapes)

    def epilogues():
        M, N, K, L = 22 * 12, 256, 128, 22
        a, b = g(M, K, seed=1, dtype=torch.bfloat16), g(N, K, seed=2, scale=0.1, dtype=torch.bfloat16)
        bias, pos = g(N, seed=3), g(L, N, seed=4)
        res = g(M, N, seed=5, dtype=torch.bfloat16)
        rs = torch.rand(M // L, device=DEV)
        base = _mm_ref(a, b) + bias
        report("gemm_tn +bias bf16 out", ops.gemm_tn(a, b, N, bias=bias), base, 1e-2)
        report("gemm_tn +bias+pos", ops.gemm_tn(a, b, N, bias=bias, pos=pos, out_dtype=torch.float32), base + pos.repeat(M // L, 1), 2e-3)
        y, pre = ops.gemm_tn(a, b, N, bias=bias, act=ops.ACT_GELU, want_preact=True, out_dtype=torch.float32)
        report("gemm_tn gelu out", y, torch.nn.functional.gelu(base), 2e-3)
        report("gemm_tn gelu preact", pre, base, 2e-3)
        ref = rs.repeat_interleave(L)[:, None] * base + res.float()
        report("gemm_tn row_scale+residual", ops.gemm_tn(a, b, N, bias=bias, residual=res, row_scale=rs, row_scale_div=L,
                                                          out_dtype=torch.float32), ref, 2e-3)
        aux = g(M, N, seed=6, dtype=torch.bfloat16)
        xr = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(xr).backward(torch.ones_like(xr))
        report("gemm_tn gelu_bwd", ops.gemm_tn(a, b, N, act=ops.ACT_GELU_BWD, aux=aux, out_dtype=torch.float32), _mm_ref(a, b) * xr.grad, 2e-3)
        acc = torch.ones(M, N, device=DEV)
        ops.gemm_tn(a, b, N, residual=acc, out=acc)
        report("gemm_tn in-place accumulate", acc, _mm_ref(a, b) + 1, 2e-3)
    check(epilogues)

    def conv_taps():
        for (Nb, T, V, Cc, Co) in ((2, 8, 22, 64, 64), (3, 32, 22, 128, 128), (2, 16, 46, 128, 128), (2, 5, 7, 64, 128)):
            x = g(Nb, Cc, T, V, seed=1, dtype=torch.bfloat16)
            w = g(Co, Cc, 9, 1, seed=2, scale=0.05)
            bias = g(Co, seed=3)
            ref = torch.nn.functional.conv2d(x.float(), w.bfloat16().float(), bias, padding=(4, 0))
            tok = x.permute(0, 2, 3, 1).reshape(Nb * T * V, Cc).contiguous()
            fw, bw = ops.conv_weight_pack(w)
            y = ops.gemm_tn(tok, fw, Co, k_per_tap=Cc, taps=9, tap_row_stride=V, tap_pad=4, rows_per_batch=T * V, batches=Nb,
                            bias=bias, out_dtype=torch.float32)
            report(f"conv9x1 fwd N={Nb} T={T} V={V} C={Cc}->{Co}", y, ref.permute(0, 2, 3, 1).reshape(-1, Co), 3e-3)
            dy = g(Nb, Co, T, V, seed=4, dtype=torch.bfloat16)
            xr = x.float().requires_grad_(True)
            torch.nn.functional.conv2d(xr, w.bfloat16().float(), bias, padding=(4, 0)).backward(dy.float())
            dtok = dy.permute(0, 2, 3, 1).reshape(Nb * T * V, Co).contiguous()
            dx = ops.gemm_tn(dtok, bw, Cc, k_per_tap=Co, taps=9, tap_row_stride=V, tap_pad=4, rows_per_batch=T * V, batches=Nb,
                             out_dtype=torch.float32)
            report(f"conv9x1 dx  N={Nb} T={T} V={V}", dx, xr.grad.permute(0, 2, 3, 1).reshape(-1, Cc), 3e-3)
    check(conv_taps)

    def fp32_split():
        M, N, K = 500, 256, 128
        a, b = g(M, K, seed=1), g(N, K, seed=2, scale=0.1)
        c = ops.gemm_tn(ops.split3(a, 0), ops.split3(b, 1), N, out_dtype=torch.float32)
        report("gemm_tn split3 fp32 parity", c, a.double() @ b.double().t(), 2e-5)
    check(fp32_split)


def grp_gemm_mn():
    def probe():
        # dx = g @ W with W [N=64 (contraction), K=64]; W = identity -> out == g
        gmat = g(128, 64, seed=1, dtype=torch.bfloat16)
        w = torch.eye(64, device=DEV, dtype=torch.bfloat16)
        c = ops.gemm_tn(gmat, w, 64, b_mn_major=True, out_dtype=torch.float32)
        ok = report("gemm_tn MN-major B probe identity", c, gmat.float(), 1e-6)
        if not ok:
            good_cols = ((c - gmat.float()).abs().max(0).values < 1e-6).nonzero().flatten().tolist()
            print("   correct columns:", good_cols)
            print("   c[0,:8]", c[0, :8].tolist(), " ref", gmat[0, :8].float().tolist())
        w2 = g(64, 64, seed=3, scale=0.2, dtype=torch.bfloat16)
        c = ops.gemm_tn(gmat, w2, 64, b_mn_major=True, out_dtype=torch.float32)
        report("gemm_tn MN-major B random 128x64x64", c, gmat.float() @ w2.float(), 2e-3)
    check(probe)

    def shapes():
        for (M, N, K) in ((256, 128, 128), (300, 768, 256), (1000, 256, 768), (4096, 1024, 512), (333, 512, 1536)):
            # dy [M, N] @ W [N, K] -> [M, K]
            dy, w = g(M, N, seed=1, dtype=torch.bfloat16), g(N, K, seed=2, scale=0.1, dtype=torch.bfloat16)
            c = ops.gemm_tn(dy, w, K, b_mn_major=True, out_dtype=torch.float32)
            report(f"gemm_tn MN-major dx M={M} N={N} K={K}", c, dy.float() @ w.float(), 2e-3)
    check(shapes)


def grp_gemm_dw():
    def probe():
        # G = [I_64; 0...] rows -> dW[n1, n2] = X[n1, n2] for n1 < 64
        M = 64
        G = torch.zeros(M, 128, device=DEV, dtype=torch.bfloat16)
        G[:64, :64] = torch.eye(64, device=DEV, dtype=torch.bfloat16)
        X = g(M, 64, seed=1, dtype=torch.bfloat16)
        dW = torch.zeros(128, 64, device=DEV)
        ops.gemm_dw(G, X, dW)
        ref = G.float().t() @ X.float()
        ok = report("gemm_dw probe identity", dW, ref, 1e-6)
        if not ok:
            print("   dW[0,:8]", dW[0, :8].tolist(), " ref", ref[0, :8].tolist())
            print("   rows ok:", ((dW - ref).abs().max(1).values < 1e-6).nonzero().flatten().tolist()[:70])
    check(probe)

    def shapes():
        for (M, N1, N2) in ((64, 128, 64), (640, 128, 128), (1000, 256, 256), (5000, 768, 256), (20000, 512, 1024), (3000, 96, 128), (2000, 16, 64)):
            G, X = g(M, N1, seed=1, scale=0.1, dtype=torch.bfloat16), g(M, N2, seed=2, dtype=torch.bfloat16)
            dW = torch.zeros(N1, N2, device=DEV)
            ops.gemm_dw(G, X, dW)
            report(f"gemm_dw M={M} N1={N1} N2={N2}", dW, G.float().t() @ X.float(), 2e-3)
        G, X = g(3000, 384, seed=3, scale=0.1, dtype=torch.bfloat16), g(3000, 192, seed=4, dtype=torch.bfloat16)
        dW = torch.zeros(128, 64, device=DEV)
        ops.gemm_dw(G, X, dW, N1=128, N2=64, ld1=64, g_col0=128, x_col0=64)
        report("gemm_dw column windows", dW, G[:, 128:256].float().t() @ X[:, 64:128].float(), 2e-3)
    check(shapes)

    def conv_dw():
        Nb, T, V, Cc, Co = 3, 16, 22, 128, 128
        x = g(Nb, Cc, T, V, seed=1, dtype=torch.bfloat16)
        dy = g(Nb, Co, T, V, seed=2, scale=0.1, dtype=torch.bfloat16)
        w = torch.zeros(Co, Cc, 9, 1, device=DEV, requires_grad=True)
        torch.nn.functional.conv2d(x.float(), w, None, padding=(4, 0)).backward(dy.float())
        tok = x.permute(0, 2, 3, 1).reshape(-1, Cc).contiguous()
        dtok = dy.permute(0, 2, 3, 1).reshape(-1, Co).contiguous()
        dW = torch.zeros(Co, Cc, 9, device=DEV)
        flat = dW.view(-1)
        for tap in range(9):
            ops.gemm_dw(dtok, tok, flat[tap:], N1=Co, N2=Cc, rows_per_batch=T * V, batches=Nb, ld1=Cc * 9, ld2=9, x_row_shift=(tap - 4) * V)
        report("conv9x1 dW via 9 shifted gemm_dw", dW, w.grad[..., 0], 3e-3)
    check(conv_dw)


# --------------------------------------------------------------------------------------------
def _attn_ref(qkv, B, L, H):
    D = qkv.shape[1] // 3
    dh = D // H
    q, k, v = qkv.view(B, L, 3, H, dh).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1)
    return (att @ v).transpose(1, 2).reshape(B * L, D)


def grp_attention():
    def attn():
        for (B, L, H, dh) in ((5, 22, 8, 32), (3, 32, 8, 64), (2, 46, 8, 32), (2, 64, 8, 64), (4, 8, 8, 16)):
            for dtp, tol in ((torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)):
                D = H * dh
                qkv = g(B * L, 3 * D, seed=L, dtype=dtp)
                o = ops.attention_fwd(qkv, B, L, H)
                qr = qkv.float().requires_grad_(True)
                oref = _attn_ref(qr, B, L, H)
                report(f"attention_fwd B={B} L={L} dh={dh} {dtp}", o, oref.detach(), tol)
                do = g(B * L, D, seed=9, dtype=dtp)
                oref.backward(do.float())
                report(f"attention_bwd B={B} L={L} dh={dh} {dtp}", ops.attention_bwd(qkv, do, B, L, H), qr.grad, tol)
    check(attn)


# --------------------------------------------------------------------------------------------
def _oracle_setup():
    from oracle import altformer_oracle as O
    return O


def _load_agcn(mod, st, prefix=""):
    sd = {k[len(prefix):]: v for k, v in st.items() if k.startswith(prefix)}
    mod.load_state_dict(sd, strict=True)


def _grad_report(tag, mod, ref_params, prefix, tol, abs_floor=1e-3):
    named = dict(mod.named_parameters())
    worst = 0.0
    for k, p in named.items():
        rg = ref_params[prefix + k].grad
        if rg is None:
            continue
        got = p.grad
        if got is None:
            RESULTS.append((f"{tag} grad {k}", False))
            print(f"FAIL {tag} grad {k}: missing")
            continue
        rn = float(rg.norm())
        if rn < 1e-5:  # analytically zero gradients: absolute check
            ok = float(got.float().norm()) < abs_floor
            RESULTS.append((f"{tag} grad {k}", ok))
            if not ok:
                print(f"FAIL {tag} grad {k}: expected ~0, got norm {float(got.float().norm()):.3e}")
            continue
        e_inf, e_l2 = rel(got.cpu(), rg)
        worst = max(worst, e_l2)
        ok = e_inf <= tol and e_l2 <= tol
        RESULTS.append((f"{tag} grad {k}", ok))
        if not ok:
            print(f"FAIL {tag} grad {k:40s} rel_inf={e_inf:.3e} rel_l2={e_l2:.3e} tol={tol:g}")
    print(f"     {tag}: worst grad rel_l2 = {worst:.3e}", flush=True)


def _oracle_run(fn, st, x, need_dx):
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in st.items()}
    xr = x.clone().requires_grad_(need_dx)
    y = fn(xr, params)
    cot = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
    (y * cot).sum().backward()
    return y.detach(), xr.grad, params, cot


def grp_gcn0():
    O = _oracle_setup()

    def gcn0_case(N, T, V, training, mode, tol, seed):
        AF.set_precision(mode)
        A = O.spatial_graph(V)
        st = O.random_state(O.agcn_spec("", 3, 128, V), seed)
        x, _ = O.synthetic_batch(N, T, V, 14, seed + 1)
        xc = x.permute(0, 3, 1, 2).contiguous()
        yr, _, params, cot = _oracle_run(lambda x_, p: O.agcn_forward(x_, p, "", A, training), st, xc, False)
        mod = ab.unit_agcn(3, 128, A).to(DEV)
        _load_agcn(mod, st)
        mod.train(training)
        y = mod(x.to(DEV).permute(0, 3, 1, 2))
        tag = f"gcn0 N={N} T={T} V={V} train={training} {mode}"
        report(tag + " fwd", y.float().cpu(), yr, tol)
        if training:
            (y.float() * cot.to(DEV)).sum().backward()
            _grad_report(tag, mod, params, "", 5 * tol)
            report(tag + " running_mean", mod.bn.running_mean.cpu(), params["bn.running_mean"], 1e-4)
            report(tag + " running_var", mod.bn.running_var.cpu(), params["bn.running_var"], 1e-4)
            report(tag + " down running_var", mod.down[1].running_var.cpu(), params["down.1.running_var"], 1e-4)
        AF.set_precision("bf16")

    check(lambda: gcn0_case(4, 8, 22, True, "fp32", 1e-4, 11))
    check(lambda: gcn0_case(4, 8, 22, False, "fp32", 1e-4, 11))
    check(lambda: gcn0_case(4, 8, 22, True, "bf16", 1e-2, 11))
    check(lambda: gcn0_case(3, 6, 46, True, "fp32", 1e-4, 12))
    check(lambda: gcn0_case(3, 6, 46, True, "bf16", 1e-2, 12))
    check(lambda: gcn0_case(32, 32, 22, True, "bf16", 1e-2, 13))
    check(lambda: gcn0_case(8, 64, 46, True, "bf16", 1e-2, 14))
    check(lambda: gcn0_case(32, 32, 22, False, "bf16", 1e-2, 13))


def grp_modules():
    O = _oracle_setup()

    def unit2d_case(Cc, N, T, V, training, mode, tol):
        AF.set_precision(mode)
        st = O.random_state(O.unit2d_spec("", Cc, Cc, 9), 21)
        x = torch.randn(N, Cc, T, V, generator=torch.Generator().manual_seed(5))
        yr, dxr, params, cot = _oracle_run(lambda x_, p: O.unit2d_forward(x_, p, "", training), st, x, True)
        mod = ab.Unit2D(Cc, Cc, 9).to(DEV)
        mod.load_state_dict(st)
        mod.train(training)
        xg = x.to(DEV).requires_grad_(True)
        y = mod(xg)
        tag = f"Unit2D C={Cc} N={N} T={T} V={V} train={training} {mode}"
        report(tag + " fwd", y.float().cpu(), yr, tol)
        (y.float() * cot.to(DEV)).sum().backward()
        report(tag + " dx", xg.grad.cpu(), dxr, 5 * tol)
        _grad_report(tag, mod, params, "", 5 * tol, abs_floor=0.05 if mode == "bf16" else 1e-3)
        AF.set_precision("bf16")

    check(lambda: unit2d_case(64, 2, 12, 22, True, "fp32", 1e-4))
    check(lambda: unit2d_case(128, 3, 32, 22, True, "bf16", 1e-2))
    check(lambda: unit2d_case(128, 2, 16, 46, False, "bf16", 1e-2))

    def block_case(D, B, L, mode, tol):
        AF.set_precision(mode)
        st = O.random_state(O.block_spec("", D), 41)
        x = torch.randn(B, L, D, generator=torch.Generator().manual_seed(6))
        yr, dxr, params, cot = _oracle_run(lambda x_, p: O.block_forward(x_, p, ""), st, x, True)
        mod = ab.Block(D, 8, mlp_ratio=2.0, qkv_bias=True, norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6)).to(DEV)
        mod.load_state_dict(st)
        xg = x.to(DEV).requires_grad_(True)
        y = mod(xg)
        tag = f"Block D={D} B={B} L={L} {mode}"
        report(tag + " fwd", y.float().cpu(), yr, tol)
        (y.float() * cot.to(DEV)).sum().backward()
        report(tag + " dx", xg.grad.cpu(), dxr, 2 * tol)
        _grad_report(tag, mod, params, "", 3 * tol)
        # sub-modules standalone
        a = mod.attn(xg.detach())
        ar = O.attention_forward(x, st, "attn.")
        report(tag + " Attention standalone", a.float().cpu(), ar, tol)
        m = mod.mlp(xg.detach())
        report(tag + " Mlp standalone", m.float().cpu(), O.mlp_forward(x, st, "mlp."), tol)
        AF.set_precision("bf16")

    check(lambda: block_case(256, 24, 22, "fp32", 1e-4))
    check(lambda: block_case(256, 24, 22, "bf16", 1.5e-2))
    check(lambda: block_case(512, 6, 32, "bf16", 1.5e-2))

    def droppath_case():
        AF.set_precision("fp32")
        D, B, L = 256, 12, 22
        st = O.random_state(O.block_spec("", D), 43)
        x = torch.randn(B, L, D, generator=torch.Generator().manual_seed(8))
        k1 = (torch.rand(B) > 0.3).float() / 0.7
        k2 = (torch.rand(B) > 0.3).float() / 0.7
        yr, dxr, params, cot = _oracle_run(lambda x_, p: O.block_forward(x_, p, "", keep=(k1, k2)), st, x, True)
        mod = ab.Block(D, 8, mlp_ratio=2.0, qkv_bias=True, drop_path=0.3, norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6)).to(DEV)
        mod.load_state_dict(st)
        calls = [k1.to(DEV), k2.to(DEV)]
        mod.drop_path.row_scale = lambda B_, dev: calls.pop(0)
        xg = x.to(DEV).requires_grad_(True)
        y = mod(xg)
        report("Block DropPath pinned masks fwd fp32", y.float().cpu(), yr, 1e-4)
        (y.float() * cot.to(DEV)).sum().backward()
        report("Block DropPath pinned masks dx fp32", xg.grad.cpu(), dxr, 2e-4)
        _grad_report("Block DropPath", mod, params, "", 3e-4)
        AF.set_precision("bf16")
    check(droppath_case)

    def agcn_case(cin, cout, N, T, V, mode, tol):
        AF.set_precision(mode)
        A = O.spatial_graph(V)
        st = O.random_state(O.agcn_spec("", cin, cout, V), 13)
        x = 0.5 * torch.randn(N, cin, T, V, generator=torch.Generator().manual_seed(9))
        yr, dxr, params, cot = _oracle_run(lambda x_, p: O.agcn_forward(x_, p, "", A, True), st, x, True)
        mod = ab.unit_agcn(cin, cout, A).to(DEV)
        mod.load_state_dict(st)
        xg = x.to(DEV).requires_grad_(True)
        y = mod(xg)
        tag = f"unit_agcn {cin}->{cout} N={N} T={T} V={V} {mode}"
        report(tag + " fwd", y.float().cpu(), yr, tol)
        (y.float() * cot.to(DEV)).sum().backward()
        report(tag + " dx", xg.grad.cpu(), dxr, 5 * tol)
        _grad_report(tag, mod, params, "", 5 * tol, abs_floor=0.05 if mode == "bf16" else 1e-3)
        AF.set_precision("bf16")

    check(lambda: agcn_case(64, 64, 2, 8, 22, "fp32", 1e-4))
    check(lambda: agcn_case(64, 128, 2, 8, 22, "fp32", 1e-4))
    check(lambda: agcn_case(128, 128, 4, 32, 22, "bf16", 1.5e-2))

    def tcn_gcn_case(mode, tol):
        AF.set_precision(mode)
        Cc, N, T, V = 64, 2, 8, 22
        A = O.spatial_graph(V)
        spec = O.OrderedDict()
        spec.update(O.agcn_spec("gcn1.", Cc, Cc, V))
        spec.update(O.unit2d_spec("tcn1.", Cc, Cc, 9))
        st = O.random_state(spec, 31)
        x = torch.randn(N, Cc, T, V, generator=torch.Generator().manual_seed(131))
        yr, dxr, params, cot = _oracle_run(lambda x_, p: O.tcn_gcn_forward(x_, p, "", A, True), st, x, True)
        mod = ab.TCN_GCN_unit(Cc, Cc, A, dropout=0.0).to(DEV)
        mod.load_state_dict(st)
        xg = x.to(DEV).requires_grad_(True)
        y = mod(xg)
        report(f"TCN_GCN_unit {mode} fwd", y.float().cpu(), yr, tol)
        (y.float() * cot.to(DEV)).sum().backward()
        report(f"TCN_GCN_unit {mode} dx", xg.grad.cpu(), dxr, 5 * tol)
        _grad_report(f"TCN_GCN_unit {mode}", mod, params, "", 5 * tol, abs_floor=0.05 if mode == "bf16" else 1e-3)
        AF.set_precision("bf16")
    check(lambda: tcn_gcn_case("fp32", 1e-4))
    check(lambda: tcn_gcn_case("bf16", 2e-2))


def grp_model():
    O = _oracle_setup()

    def model_case(style, N, T, V, cls, mode, tol, training=True):
        AF.set_precision(mode)
        A = O.spatial_graph(V)
        st = O.random_state(O.model_spec(3, cls, T, V), 61)
        x, _ = O.synthetic_batch(N, T, V, cls, 161)
        yr, _, params, cot = _oracle_run(lambda x_, p: O.model_forward(x_, p, A, style, training), st, x, False)
        graph = "graph.SHRE" if V == 22 else "graph.LMDHG"
        mod = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style=style, graph=graph, graph_args={"labeling_mode": "spatial"})
        mod.load_state_dict(st)
        mod = mod.to(DEV)
        for m in mod.modules():
            if type(m).__name__ == "DropPath":
                m.drop_prob = 0.0
        mod.train(training)
        y = mod(x.to(DEV))
        tag = f"model style={style} N={N} T={T} V={V} {mode} train={training}"
        report(tag + " logits", y.float().cpu(), yr, tol)
        if training:
            (y.float() * cot.to(DEV)).sum().backward()
            _grad_report(tag, mod, params, "", 6 * tol, abs_floor=0.1 if mode == "bf16" else 1e-3)
        AF.set_precision("bf16")

    check(lambda: model_case("ST", 2, 8, 22, 14, "fp32", 2e-4))
    check(lambda: model_case("TS", 2, 8, 22, 14, "fp32", 2e-4))
    check(lambda: model_case(None, 2, 8, 22, 28, "fp32", 2e-4))
    check(lambda: model_case("ST", 4, 32, 22, 28, "bf16", 3e-2))
    check(lambda: model_case("TS", 4, 32, 22, 28, "bf16", 3e-2))
    check(lambda: model_case("ST", 2, 16, 46, 14, "bf16", 3e-2, training=False))


def grp_trainer():
    def train_steps():
        from oracle import altformer_oracle as O
        torch.manual_seed(0)
        N, T, V, cls = 16, 32, 22, 28
        x, yl = O.synthetic_batch(N, T, V, cls, 1234)
        x, yl = x.to(DEV), yl.to(DEV)
        losses = {}
        for use_graph in (False, True):
            torch.manual_seed(0)
            mod = ab.ST_GCN_AltFormer(3, cls, num_frame=T, num_joints=V, style="ST", graph="graph.SHRE", graph_args={"labeling_mode": "spatial"}).to(DEV)
            for m in mod.modules():
                if type(m).__name__ == "DropPath":
                    m.drop_prob = 0.0
            tr = ab.DataParallelTrainer(mod, use_graph=use_graph)
            ls = []
            for it in range(6):
                loss, _ = tr.step(x, yl)
                ls.append(float(loss))
            losses[use_graph] = ls
            print("   losses graph=%s:" % use_graph, ["%.4f" % v for v in ls], flush=True)
        ok = losses[False][-1] < losses[False][0]
        RESULTS.append(("trainer loss decreases", ok))
        print(("PASS" if ok else "FAIL") + " trainer loss decreases")
        report("trainer graph == eager losses", torch.tensor(losses[True]), torch.tensor(losses[False]), 2e-2)
    check(train_steps)


GROUPS = {"simt": grp_simt, "elementwise": grp_elementwise, "gemm_tn": grp_gemm_tn, "gemm_mn": grp_gemm_mn, "gemm_dw": grp_gemm_dw,
          "attention": grp_attention, "gcn0": grp_gcn0, "modules": grp_modules, "model": grp_model, "trainer": grp_trainer}

if __name__ == "__main__":
    names = sys.argv[1:] or list(GROUPS)
    print("device:", torch.cuda.get_device_name(0), "| lib version", ab._lib.lib().afb_version(), flush=True)
    for n in names:
        print(f"===== group {n} =====", flush=True)
        GROUPS[n]()
    npass = sum(1 for _, ok in RESULTS if ok)
    print(f"SUMMARY {' '.join(names)}: {npass}/{len(RESULTS)} passed", flush=True)
    for name, ok in RESULTS:
        if not ok:
            print("  failed:", name)
    sys.exit(0 if npass == len(RESULTS) else 1)
