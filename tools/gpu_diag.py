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
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
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
        report("adamw step counter", step.float().cpu(), torch.tensor(4.0), 0)
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

    def fast_epilogues():
        # bf16 outputs take the specialised straight-line epilogues (EPI_* kinds); BN 64 / 128 / 256 tiles,
        # ragged last m-tile, DropPath row scales in both modes
        L = 22
        for (M, N, K) in ((22 * 13, 64, 128), (22 * 50, 128, 256), (22 * 300, 256, 256), (22 * 40, 512, 256), (22 * 400, 256, 512)):
            a, b = g(M, K, seed=1, dtype=torch.bfloat16), g(N, K, seed=2, scale=0.1, dtype=torch.bfloat16)
            bias = g(N, seed=3)
            res = g(M, N, seed=5, dtype=torch.bfloat16)
            rs = (torch.rand(M // L, device=DEV) > 0.3).float() / 0.7
            rsr = rs.repeat_interleave(L)[:, None]
            mm = _mm_ref(a, b)
            tag = f"{M}x{N}x{K}"
            report(f"fast bias {tag}", ops.gemm_tn(a, b, N, bias=bias), mm + bias, 1e-2)
            report(f"fast bias+res {tag}", ops.gemm_tn(a, b, N, bias=bias, residual=res), mm + bias + res.float(), 1e-2)
            report(f"fast bias+res+rs {tag}", ops.gemm_tn(a, b, N, bias=bias, residual=res, row_scale=rs, row_scale_div=L),
                   rsr * (mm + bias) + res.float(), 1e-2)
            report(f"fast bias+res+rs(bias only) {tag}",
                   ops.gemm_tn(a, b, N, bias=bias, residual=res, row_scale=rs, row_scale_div=L, row_scale_bias_only=True),
                   mm + rsr * bias + res.float(), 1e-2)
            y, pre = ops.gemm_tn(a, b, N, bias=bias, act=ops.ACT_GELU, want_preact=True)
            report(f"fast gelu {tag}", y, torch.nn.functional.gelu(mm + bias), 1e-2)
            report(f"fast gelu preact {tag}", pre, mm + bias, 1e-2)
            y, pre = ops.gemm_tn(a, b, N, bias=bias, act=ops.ACT_GELU, want_preact=True, row_scale=rs, row_scale_div=L)
            report(f"fast gelu*rs {tag}", y, rsr * torch.nn.functional.gelu(mm + bias), 1e-2)
            report(f"fast gelu*rs preact {tag}", pre, mm + bias, 1e-2)
            aux = g(M, N, seed=6, dtype=torch.bfloat16)
            xr = aux.float().requires_grad_(True)
            torch.nn.functional.gelu(xr).backward(torch.ones_like(xr))
            report(f"fast gelu_bwd {tag}", ops.gemm_tn(a, b, N, act=ops.ACT_GELU_BWD, aux=aux), mm * xr.grad, 1.2e-2)
            report(f"fast gelu_bwd*rs {tag}", ops.gemm_tn(a, b, N, act=ops.ACT_GELU_BWD, aux=aux, row_scale=rs, row_scale_div=L),
                   rsr * mm * xr.grad, 1.2e-2)
            report(f"fast plain*rs {tag}", ops.gemm_tn(a, b, N, row_scale=rs, row_scale_div=L), rsr * mm, 1e-2)
    check(fast_epilogues)

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
        for (M, N1, N2) in ((1000, 256, 256), (5000, 768, 256), (700, 512, 1024), (900, 96, 128), (30000, 256, 128), (2000, 256, 64),
                           (180224, 512, 256)):   # fused bias gradient
            G, X = g(M, N1, seed=5, scale=0.1, dtype=torch.bfloat16), g(M, N2, seed=6, dtype=torch.bfloat16)
            dW, db = torch.zeros(N1, N2, device=DEV), torch.zeros(N1, device=DEV)
            ops.gemm_dw(G, X, dW, dbias=db)
            report(f"gemm_dw+dbias M={M} N1={N1} N2={N2} dW", dW, G.float().t() @ X.float(), 2e-3)
            report(f"gemm_dw+dbias M={M} N1={N1} N2={N2} dbias", db, G.float().sum(0), 2e-3)
        for (M, N1, N2, L) in ((22 * 100, 256, 256, 22), (32 * 300, 512, 1024, 32), (22 * 77, 256, 512, 22)):   # row-scaled bias gradient
            G, X = g(M, N1, seed=7, scale=0.1, dtype=torch.bfloat16), g(M, N2, seed=8, dtype=torch.bfloat16)
            rs = (torch.rand(M // L, device=DEV) > 0.3).float() / 0.7
            dW, db = torch.zeros(N1, N2, device=DEV), torch.zeros(N1, device=DEV)
            ops.gemm_dw(G, X, dW, dbias=db, dbias_row_scale=rs, row_scale_div=L)
            report(f"gemm_dw+dbias*rs M={M} N1={N1} N2={N2} dW", dW, G.float().t() @ X.float(), 2e-3)
            report(f"gemm_dw+dbias*rs M={M} N1={N1} N2={N2} dbias", db, (rs.repeat_interleave(L)[:, None] * G.float()).sum(0), 2e-3)
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
        # three taps per launch (shared dY tile), tap-major scratch, bias gradient from the first group
        tmp = torch.zeros(9, Co, Cc, device=DEV)
        db = torch.zeros(Co, device=DEV)
        for t0 in range(0, 9, 3):
            ops.gemm_dw(dtok, tok, tmp.view(-1)[t0 * Co * Cc:], N1=Co, N2=Cc, rows_per_batch=T * V, batches=Nb, ld1=Cc, ld2=1,
                        x_row_shift=(t0 - 4) * V, taps=3, tap_row_stride=V, tap_dw_stride=Co * Cc, dbias=db if t0 == 0 else None)
        dW2 = torch.full((Co, Cc, 9), 0.5, device=DEV)
        ops.conv_dw_unpack(tmp, dW2, Co, Cc, 9)
        report("conv9x1 dW via 3 x 3-tap gemm_dw + unpack", dW2 - 0.5, w.grad[..., 0], 3e-3)
        report("conv9x1 dbias from the first tap group", db, dtok.float().sum(0), 3e-3)
        tmp2 = torch.zeros(2, Co, Cc, device=DEV)
        ops.gemm_dw(dtok, tok, tmp2.view(-1), N1=Co, N2=Cc, rows_per_batch=T * V, batches=Nb, ld1=Cc, ld2=1, x_row_shift=-V, taps=2,
                    tap_row_stride=V, tap_dw_stride=Co * Cc)
        report("gemm_dw taps=2", tmp2, w.grad[..., 0].permute(2, 0, 1)[3:5], 3e-3)
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
        for (B, L, H, dh) in ((5, 22, 8, 32), (3, 32, 8, 64), (2, 46, 8, 32), (2, 64, 8, 64), (4, 8, 8, 16),
                              (4, 32, 8, 32), (3, 17, 8, 32), (2, 25, 4, 32), (3, 24, 8, 32), (2, 16, 8, 32), (2, 9, 4, 64),
                              (2, 100, 8, 32), (1, 180, 8, 64), (2, 65, 4, 16), (1, 256, 2, 32)):   # L > 64: coverage kernels
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
                keep = torch.tensor([0.0 if i % 3 == 1 else 1.25 for i in range(B)], device=DEV)
                report(f"attention_fwd*keep B={B} L={L} dh={dh} {dtp}", ops.attention_fwd(qkv, B, L, H, out_scale=keep),
                       keep.repeat_interleave(L)[:, None] * oref.detach(), tol)

    def attn_variants():
        """every forward kernel on the shapes all of them cover (AFB_ATTN_TC: 0 = warp-level MMA, 1 = tcgen05 with the
        probabilities in tensor memory, 2 = tcgen05 with the probabilities staged in shared memory); multi-tile batches,
        partial last tiles, every (LP, dh) instantiation incl. the runtime-length ones"""
        saved = os.environ.get("AFB_ATTN_TC")
        try:
            for (B, L, H, dh) in ((5, 22, 8, 32), (1237, 22, 8, 32), (301, 32, 8, 32), (7, 9, 4, 32), (130, 46, 8, 32),
                                  (75, 64, 8, 32), (3, 50, 2, 32), (203, 32, 8, 64), (9, 13, 4, 64), (41, 46, 8, 64),
                                  (66, 64, 4, 64), (5, 40, 3, 64), (1, 1, 2, 32), (2, 33, 6, 32), (1, 64, 2, 32), (1, 2, 2, 64)):
                D = H * dh
                qkv = g(B * L, 3 * D, seed=L + B, dtype=torch.bfloat16)
                oref = _attn_ref(qkv.float(), B, L, H)
                keep = torch.tensor([0.0 if i % 3 == 1 else 1.25 for i in range(B)], device=DEV)
                for mode, name in (("0", "mma"), ("1", "tc"), ("2", "tc-smemP")):
                    os.environ["AFB_ATTN_TC"] = mode
                    report(f"attention_fwd[{name}] B={B} L={L} H={H} dh={dh}", ops.attention_fwd(qkv, B, L, H), oref, 1e-2)
                    report(f"attention_fwd[{name}]*keep B={B} L={L} H={H} dh={dh}", ops.attention_fwd(qkv, B, L, H, out_scale=keep),
                           keep.repeat_interleave(L)[:, None] * oref, 1e-2)
            # backward: warp-level kernel vs tcgen05 kernel (dh 32) vs torch autograd
            for (B, L, H, dh) in ((5, 22, 8, 32), (1237, 22, 8, 32), (301, 32, 8, 32), (7, 9, 4, 32), (130, 46, 8, 32), (75, 64, 8, 32),
                                  (3, 50, 2, 32), (1, 1, 2, 32), (2, 33, 6, 32), (1, 64, 2, 32)):
                D = H * dh
                qkv = g(B * L, 3 * D, seed=L + B, dtype=torch.bfloat16)
                do = g(B * L, D, seed=5, dtype=torch.bfloat16)
                qr = qkv.float().requires_grad_(True)
                _attn_ref(qr, B, L, H).backward(do.float())
                for mode, name in (("0", "mma"), ("1", "tc")):
                    os.environ["AFB_ATTN_TC_BWD"] = mode
                    got = ops.attention_bwd(qkv, do, B, L, H)
                    for k, part in enumerate(("dq", "dk", "dv")):
                        report(f"attention_bwd[{name}] {part} B={B} L={L} H={H} dh={dh}", got[:, k * D:(k + 1) * D], qr.grad[:, k * D:(k + 1) * D], 1e-2)
        finally:
            os.environ.pop("AFB_ATTN_TC_BWD", None)
            if saved is None:
                os.environ.pop("AFB_ATTN_TC", None)
            else:
                os.environ["AFB_ATTN_TC"] = saved
    check(attn)
    check(attn_variants)


# --------------------------------------------------------------------------------------------
def _module_groups():
    from tools import gpu_diag_modules
    return gpu_diag_modules.GROUPS


GROUPS = {"simt": grp_simt, "elementwise": grp_elementwise, "gemm_tn": grp_gemm_tn, "gemm_mn": grp_gemm_mn, "gemm_dw": grp_gemm_dw,
          "attention": grp_attention}
MODULE_GROUPS = ("gcn0", "modules", "model", "trainer")


def run_group(name):
    if name in GROUPS:
        return GROUPS[name]()
    return _module_groups()[name]()

if __name__ == "__main__":
    sys.modules.setdefault("tools.gpu_diag", sys.modules["__main__"])   # one shared RESULTS list
    names = sys.argv[1:] or list(GROUPS) + list(MODULE_GROUPS)
    print("device:", torch.cuda.get_device_name(0), "| lib version", ab._lib.lib().afb_version(), flush=True)
    for n in names:
        print(f"===== group {n} =====", flush=True)
        run_group(n)
    npass = sum(1 for _, ok in RESULTS if ok)
    print(f"SUMMARY {' '.join(names)}: {npass}/{len(RESULTS)} passed", flush=True)
    for name, ok in RESULTS:
        if not ok:
            print("  failed:", name)
    sys.exit(0 if npass == len(RESULTS) else 1)
