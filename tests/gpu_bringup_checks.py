"""GPU bring-up checks, each run in its own subprocess with a timeout so that a hung kernel cannot take the
whole call down.  Usage (on a GPU box):  python tests/gpu_bringup_checks.py [name ...]   (no names = all)
Results are appended to gpurun_out/gpu_check.log.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def _mm_ref_s32(A, W):
    import torch
    return (A.cpu().to(torch.int32) @ W.cpu().to(torch.int32).T).to(A.device)


def _gemm_i8_cases(cg):
    import torch
    from llamax_b200 import ops
    ops.set_gemm_cta_group(cg)
    torch.manual_seed(0)
    ok = True
    for (M, N, K) in [(128, 256, 128), (256, 512, 512), (300, 264, 208), (1024, 1024, 4096), (2048, 14336, 4096)]:
        A = torch.randint(-127, 128, (M, K), device="cuda", dtype=torch.int8)
        W = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8)
        out = ops.int8_gemm_s32(A, W)
        torch.cuda.synchronize()
        if M * N * K <= 2**31:
            ref = _mm_ref_s32(A, W)
        else:
            ref = torch._int_mm(A, W.T)
        bad = (out != ref)
        nbad = int(bad.sum())
        print(f"  int8 s32 cg={cg} M={M} N={N} K={K}: mismatches={nbad}/{M*N}", flush=True)
        if nbad:
            ok = False
            idx = bad.nonzero()[:5].tolist()
            print("    first bad idx:", idx, "got", [int(out[i, j]) for i, j in idx], "ref", [int(ref[i, j]) for i, j in idx])
            rows_bad = bad.any(1).sum().item(); cols_bad = bad.any(0).sum().item()
            print(f"    rows with errors {rows_bad}/{M}, cols with errors {cols_bad}/{N}")
    # dequant epilogue, bit-exact vs restated reference epilogue
    M, N, K = 512, 768, 1024
    A = torch.randint(-127, 128, (M, K), device="cuda", dtype=torch.int8)
    W = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8)
    sa = (torch.rand(M, device="cuda") * 0.1).bfloat16()
    sw = (torch.rand(N, device="cuda") * 0.01).bfloat16()
    out = ops.int8_gemm_dequant(A, W, sa, sw)
    ref = ((_mm_ref_s32(A, W).float() * sa.float()[:, None]) * sw.float()[None, :]).bfloat16()
    nbad = int((out != ref).sum())
    print(f"  int8 dequant cg={cg}: bf16 mismatches={nbad}/{M*N}", flush=True)
    ok &= nbad == 0
    # + lora + residual (tolerance)
    R = 8
    h = torch.randn(M, R, device="cuda").bfloat16()
    lb = (torch.randn(N, R, device="cuda") * 0.1).bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16()
    out = ops.int8_gemm_dequant(A, W, sa, sw, lora_h=h, lora_b=lb, lora_scale=2.0, resid=res)
    ref32 = (_mm_ref_s32(A, W).float() * sa.float()[:, None]) * sw.float()[None, :] + 2.0 * (h.float() @ lb.float().T) + res.float()
    err = (out.float() - ref32).abs().max().item() / ref32.abs().max().item()
    print(f"  int8 dequant+lora+resid cg={cg}: max rel-to-max err={err:.3e}", flush=True)
    ok &= err < 5e-3
    return ok


@check
def gemm_i8_cg1():
    return _gemm_i8_cases(1)


@check
def gemm_i8_cg2():
    return _gemm_i8_cases(2)


def _gemm_bf16_cases(cg):
    import torch
    from llamax_b200 import ops
    ops.set_gemm_cta_group(cg)
    torch.manual_seed(1)
    ok = True
    for (M, N, K) in [(128, 256, 64), (256, 512, 512), (300, 264, 200), (2048, 4096, 4096), (1000, 16, 4096)]:
        A = torch.randn(M, K, device="cuda").bfloat16()
        B = torch.randn(N, K, device="cuda").bfloat16()
        out = ops.bf16_gemm(A, B)
        ref = A.float() @ B.float().T
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"  bf16 cg={cg} M={M} N={N} K={K}: max rel-to-max err={err:.3e}", flush=True)
        ok &= err < 5e-3
    return ok


@check
def gemm_bf16_cg1():
    return _gemm_bf16_cases(1)


@check
def gemm_bf16_cg2():
    return _gemm_bf16_cases(2)


@check
def gemm_perf():
    import torch
    from llamax_b200 import ops
    ok = True
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.int8)
    for cg in (1, 2):
        ops.set_gemm_cta_group(cg)
        for (M, N, K) in [(8192, 8192, 8192), (16384, 4096, 4096), (16384, 14336, 4096), (16384, 4096, 14336)]:
            for kind in ("i8", "bf16"):
                if kind == "i8":
                    A = torch.randint(-127, 128, (M, K), device="cuda", dtype=torch.int8)
                    W = torch.randint(-127, 128, (N, K), device="cuda", dtype=torch.int8)
                    sa = torch.rand(M, device="cuda").bfloat16(); sw = torch.rand(N, device="cuda").bfloat16()
                    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
                    f = lambda: ops.int8_gemm_dequant(A, W, sa, sw, out=out)
                    g = lambda: torch._int_mm(A, W.T)
                else:
                    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
                    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
                    f = lambda: ops.bf16_gemm(A, W, out=out)
                    g = lambda: torch.matmul(A, W.T)
                res = []
                for fn in (f, g):
                    for _ in range(3): fn()
                    ts = []
                    for _ in range(5):
                        flush.zero_()
                        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    res.append(min(ts))
                tf = 2 * M * N * K / 1e9
                print(f"  perf cg={cg} {kind} M={M} N={N} K={K}: ours {res[0]:.3f} ms = {tf/res[0]:.0f} T/s | torch {res[1]:.3f} ms = {tf/res[1]:.0f} T/s", flush=True)
    return ok


@check
def elementwise():
    import torch
    from llamax_b200 import ops
    from oracle import ref_ops as R
    torch.manual_seed(2)
    ok = True
    for (M, K) in [(64, 512), (300, 4096), (128, 14336), (5, 1792)]:
        x = (torch.randn(M, K) * 3).bfloat16()
        x[0].zero_()
        q_ref, s_ref = R.quantize_int8_rowwise(x)
        q, s = ops.rowquant_int8(x.cuda())
        bq = int((q.cpu() != q_ref).sum()); bs = int((s.cpu() != s_ref).sum())
        print(f"  rowquant M={M} K={K}: code mismatches={bq} scale mismatches={bs}", flush=True)
        ok &= bq == 0 and bs == 0
    for (M, D) in [(64, 512), (300, 4096)]:
        x = torch.randn(M, D).bfloat16(); w = (1 + 0.1 * torch.randn(D)).bfloat16()
        y_ref = R.rmsnorm_ref(x, w)
        y, rstd, q8, qs = ops.rmsnorm_fwd(x.cuda(), w.cuda(), 1e-5, quant=True)
        nb = int((y.cpu() != y_ref).sum())
        q_ref, s_ref = R.quantize_int8_rowwise(y.cpu())
        bq = int((q8.cpu() != q_ref).sum()); bs = int((qs.cpu() != s_ref).sum())
        print(f"  rmsnorm M={M} D={D}: bf16 mismatches={nb}/{M*D} (1-ulp rounding flips allowed), fused-quant code mismatches vs quant(y)={bq}, scale={bs}", flush=True)
        ok &= nb <= M * D * 0.002 and bq == 0 and bs == 0
        dy = torch.randn(M, D).bfloat16(); dres = torch.randn(M, D).bfloat16()
        dx_ref, dw_ref = R.rmsnorm_bwd_f32(dy, x, w)
        dx, dw = ops.rmsnorm_bwd(dy.cuda(), x.cuda(), w.cuda(), rstd, dres.cuda())
        e1 = (dx.cpu().float() - (dx_ref + dres.float())).abs().max().item() / (dx_ref + dres.float()).abs().max().item()
        e2 = (dw.cpu().float() - dw_ref).abs().max().item() / dw_ref.abs().max().item()
        print(f"  rmsnorm_bwd M={M} D={D}: dx err={e1:.3e} dw err={e2:.3e}", flush=True)
        ok &= e1 < 1e-2 and e2 < 1e-2
    for (M, Fd) in [(64, 1792), (100, 14336)]:
        ab = torch.randn(M, 2 * Fd).bfloat16()
        a, b = ab[:, :Fd], ab[:, Fd:]
        g_ref = R.swiglu_ref(a, b)
        abc = ab.cuda()
        g, q8, qs = ops.swiglu_fwd(abc[:, :Fd], abc[:, Fd:], quant=True)
        nb = int((g.cpu() != g_ref).sum())
        q_ref, s_ref = R.quantize_int8_rowwise(g.cpu())
        bq = int((q8.cpu() != q_ref).sum())
        print(f"  swiglu M={M} F={Fd}: bf16 mismatches={nb}/{M*Fd}, fused-quant mismatches={bq}", flush=True)
        ok &= nb <= M * Fd * 0.002 and bq == 0
        dg = torch.randn(M, Fd).bfloat16()
        da_ref, db_ref = R.swiglu_bwd_f32(dg, a, b)
        da, db, g2 = ops.swiglu_bwd(dg.cuda(), abc[:, :Fd], abc[:, Fd:], want_g=True)
        e1 = (da.cpu().float() - da_ref).abs().max().item() / da_ref.abs().max().item()
        e2 = (db.cpu().float() - db_ref).abs().max().item() / db_ref.abs().max().item()
        e3 = (g2.cpu().float() - g_ref.float()).abs().max().item()
        print(f"  swiglu_bwd: da err={e1:.3e} db err={e2:.3e} g err={e3:.3e}", flush=True)
        ok &= e1 < 1e-2 and e2 < 1e-2
    # rope
    B, S, H, D = 2, 300, 6, 128
    rope = R.build_rope(D, 512, 500000, True)
    x = torch.randn(B, S, H, D).bfloat16()
    y_ref = R.apply_rope(x, rope[:S])
    xc = torch.cat([x.reshape(B * S, H * D), torch.zeros(B * S, 256).bfloat16()], 1).cuda()  # extra cols untouched
    ops.rope_(xc, rope.cuda(), B, S, H, D)
    nb = int((xc[:, : H * D].cpu().view(B, S, H, D) != y_ref).sum())
    print(f"  rope: mismatches={nb}/{x.numel()} extra-cols-untouched={bool((xc[:, H*D:] == 0).all())}", flush=True)
    ok &= nb == 0
    ops.rope_(xc, rope.cuda(), B, S, H, D, inverse=True)
    e = (xc[:, : H * D].cpu().view(B, S, H, D).float() - x.float()).abs().max().item()
    print(f"  rope inverse roundtrip max abs err={e:.3e}", flush=True)
    ok &= e < 0.05
    # dequant weight
    N, K = 264, 208 + 16 * 3
    w8 = torch.randint(-127, 128, (N, K), dtype=torch.int8); s = (torch.rand(N) * 0.01).bfloat16()
    o1 = ops.dequant_weight(w8.cuda(), s.cuda(), transpose=False, apply_scale=False).cpu()
    o2 = ops.dequant_weight(w8.cuda(), s.cuda(), transpose=True, apply_scale=True).cpu()
    ok1 = bool((o1 == w8.bfloat16()).all()); ok2 = bool((o2 == (w8.float() * s.float()[:, None]).bfloat16().T).all())
    print(f"  dequant plain ok={ok1} transposed+scaled ok={ok2}", flush=True)
    ok &= ok1 and ok2
    # lora wgrad
    M, Pn, Rr = 1000, 512, 8
    X = torch.randn(M, Pn).bfloat16(); Hh = torch.randn(M, Rr).bfloat16()
    o = ops.lora_wgrad(X.cuda(), Hh.cuda(), 0.5).cpu()
    ref = 0.5 * X.float().T @ Hh.float()
    e = (o - ref).abs().max().item() / ref.abs().max().item()
    print(f"  lora_wgrad err={e:.3e}", flush=True)
    ok &= e < 1e-4
    return ok


def _attn_case(B, S, Hq, Hkv, P, do_bwd=True, seed=0):
    import torch
    from llamax_b200 import ops
    from oracle import ref_ops as R
    D = 128
    torch.manual_seed(seed)
    ld = (Hq + 2 * Hkv) * D
    qkv = torch.randn(B * S, ld).bfloat16()
    q, k, v = qkv[:, : Hq * D], qkv[:, Hq * D : (Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D :]
    dout = torch.randn(B * S, Hq * D).bfloat16()
    to4 = lambda t, H: t.reshape(B, S, H, D).transpose(1, 2)
    o_ref, dq_ref, dk_ref, dv_ref = R.attention_ref_grads(to4(q, Hq), to4(k, Hkv), to4(v, Hkv), to4(dout, Hq), P, torch.float64)
    g = qkv.cuda()
    qc, kc, vc = g[:, : Hq * D], g[:, Hq * D : (Hq + Hkv) * D], g[:, (Hq + Hkv) * D :]
    o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, P)
    torch.cuda.synchronize()
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    e_o = rel(to4(o.cpu(), Hq), o_ref)
    # lse check
    s = (to4(q, Hq).double() @ to4(k, Hkv).double().repeat_interleave(Hq // Hkv, 1).transpose(-1, -2)) / D ** 0.5
    s = s.masked_fill(~R.prefix_lm_mask(S, P), float("-inf"))
    e_l = (lse.cpu().double() - torch.logsumexp(s, -1)).abs().max().item()
    msg = f"  attn B={B} S={S} Hq={Hq} Hkv={Hkv} P={P}: out err={e_o:.3e} lse abs err={e_l:.3e}"
    ok = e_o < 1e-2 and e_l < 1e-2
    if do_bwd:
        dqkv = torch.zeros_like(g)
        dq, dk, dv = dqkv[:, : Hq * D], dqkv[:, Hq * D : (Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D :]
        ops.attn_bwd(qc, kc, vc, o, lse, dout.cuda(), dq, dk, dv, B, S, Hq, Hkv, D, P)
        torch.cuda.synchronize()
        e_q, e_k, e_v = rel(to4(dq.cpu(), Hq), dq_ref), rel(to4(dk.cpu(), Hkv), dk_ref), rel(to4(dv.cpu(), Hkv), dv_ref)
        msg += f" | dq err={e_q:.3e} dk err={e_k:.3e} dv err={e_v:.3e}"
        ok = ok and max(e_q, e_k, e_v) < 1e-2
    print(msg, flush=True)
    return ok


@check
def attn_fwd_small():
    ok = True
    for (B, S, Hq, Hkv, P) in [(1, 128, 1, 1, 0), (1, 256, 4, 1, 0), (2, 512, 8, 2, 128), (1, 300, 4, 2, 70), (1, 1628, 4, 1, 1500)]:
        ok &= _attn_case(B, S, Hq, Hkv, P, do_bwd=False)
    return ok


@check
def attn_bwd_small():
    ok = True
    for (B, S, Hq, Hkv, P) in [(1, 128, 1, 1, 0), (1, 256, 4, 1, 0), (2, 512, 8, 2, 128), (1, 300, 4, 2, 70), (1, 1628, 4, 1, 1500),
                               (1, 2048, 8, 2, 0)]:
        ok &= _attn_case(B, S, Hq, Hkv, P, do_bwd=True)
    return ok


@check
def attn_perf():
    import torch
    from llamax_b200 import ops
    import torch.nn.functional as F
    D, Hq, Hkv = 128, 32, 8
    for (B, S, P) in [(8, 2048, 0), (2, 8192, 0), (8, 2048, 1024), (9, 1756, 1500)]:
        ld = (Hq + 2 * Hkv) * D
        g = torch.randn(B * S, ld, device="cuda").bfloat16()
        qc, kc, vc = g[:, : Hq * D], g[:, Hq * D : (Hq + Hkv) * D], g[:, (Hq + Hkv) * D :]
        dout = torch.randn(B * S, Hq * D, device="cuda").bfloat16()
        dqkv = torch.empty_like(g)
        dq, dk, dv = dqkv[:, : Hq * D], dqkv[:, Hq * D : (Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D :]
        o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, P)
        pairs = S * P + (S - P) * (S - P + 1) / 2
        fl_f = 4 * B * Hq * D * pairs
        def timeit(fn, n=5):
            for _ in range(2): fn()
            ts = []
            for _ in range(n):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            return min(ts)
        tf = timeit(lambda: ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, P))
        tb = timeit(lambda: ops.attn_bwd(qc, kc, vc, o, lse, dout, dq, dk, dv, B, S, Hq, Hkv, D, P))
        msg = f"  attn perf B={B} S={S} P={P}: fwd {tf:.3f} ms = {fl_f/tf/1e9:.0f} TF/s | bwd {tb:.3f} ms = {2.5*fl_f/tb/1e9:.0f} TF/s"
        if P == 0:
            q4 = qc.reshape(B, S, Hq, D).transpose(1, 2); k4 = kc.reshape(B, S, Hkv, D).transpose(1, 2); v4 = vc.reshape(B, S, Hkv, D).transpose(1, 2)
            q4 = q4.detach().requires_grad_(True)
            ts = timeit(lambda: F.scaled_dot_product_attention(q4, k4, v4, is_causal=True, enable_gqa=True))
            msg += f" | torch SDPA fwd {ts:.3f} ms = {fl_f/ts/1e9:.0f} TF/s"
        print(msg, flush=True)
    return True


def main():
    names = sys.argv[1:] or list(CHECKS)
    if len(names) == 1 and names[0].startswith("--run="):
        name = names[0][6:]
        ok = CHECKS[name]()
        print(f"RESULT {name}: {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "a")
    summary = []
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), f"--run={name}"], capture_output=True, text=True,
                               timeout=int(os.environ.get("CHECK_TIMEOUT", "240")), cwd=ROOT)
            out, rc = r.stdout + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out = ((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")) + "\nTIMEOUT"
            rc = -9
        msg = f"==== {name}: rc={rc} ({time.time()-t0:.1f}s)\n{out}\n"
        print(msg, flush=True)
        log.write(msg); log.flush()
        summary.append((name, rc))
    print("SUMMARY", summary)
    log.write(f"SUMMARY {summary}\n")


if __name__ == "__main__":
    main()
