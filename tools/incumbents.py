"""Incumbent GPU paths of the reference on the same box, next to this repo's kernels ("the
reference GPU path each kernel must beat"). Evidence under profiles/, not bench lines.

  * attention: F.scaled_dot_product_attention forward AND backward (cuDNN and flash backends; llama.py:134-137),
    torch.compile'd flex_attention with the prefix-LM BlockMask, K/V expanded by repeat_interleave exactly as the
    reference does (llama.py:129-132), vs ops.attn_fwd / ops.attn_bwd
  * INT8 GEMM: the reference's Triton kernel (baseline/_ref/subclasses/int8_mm.py:50-118, autotuned) and cuBLASLt
    torch._int_mm vs ops.int8_gemm_dequant
  * INT8 / bf16 library peaks (8192^3) in the same process
CUDA events over graph-free loops of `reps` calls, best of 3, after warm-up; TFLOP/s over UNMASKED pairs only."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from llamax_b200 import ops

dev = "cuda"
which = set(sys.argv[1:]) or {"attn", "flex", "gemm", "peak"}


def timeit(fn, reps=10, n=3):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        best = t if best is None or t < best else best
    return best


Hq, Hkv, D = 32, 8, 128
if "attn" in which or "flex" in which:
    from torch.nn.attention import SDPBackend, sdpa_kernel

    print("== attention, Hq 32 / Hkv 8 / head_dim 128, bf16; TFLOP/s = 4 (fwd) / 10 (bwd) * B*Hq*D*unmasked_pairs / time")
    for B, S, P in ((8, 2048, 0), (8, 1756, 1500), (2, 8192, 0)):
        torch.manual_seed(0)
        ld = (Hq + 2 * Hkv) * D
        g = torch.randn(B * S, ld, device=dev).bfloat16()
        q2, k2, v2 = g[:, : Hq * D], g[:, Hq * D : (Hq + Hkv) * D], g[:, (Hq + Hkv) * D :]
        pairs = S * P + (S - P) * (S - P + 1) / 2
        fl = 4.0 * B * Hq * D * pairs
        dout2 = torch.randn(B * S, Hq * D, device=dev).bfloat16()
        dqkv = torch.empty_like(g)
        dq, dk, dv = dqkv[:, : Hq * D], dqkv[:, Hq * D : (Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D :]
        o, lse = ops.attn_fwd(q2, k2, v2, B, S, Hq, Hkv, D, P)
        tf = timeit(lambda: ops.attn_fwd(q2, k2, v2, B, S, Hq, Hkv, D, P))
        tb = timeit(lambda: ops.attn_bwd(q2, k2, v2, o, lse, dout2, dq, dk, dv, B, S, Hq, Hkv, D, P))
        print(f"B={B} S={S} P={P}  llamax_b200          fwd {tf*1e3:8.1f} us {fl/tf/1e9:7.1f} TF/s | bwd {tb*1e3:8.1f} us {2.5*fl/tb/1e9:7.1f} TF/s", flush=True)
        # library layouts: [B, H, S, D]
        q = q2.view(B, S, Hq, D).transpose(1, 2).contiguous().requires_grad_(True)
        k = k2.view(B, S, Hkv, D).transpose(1, 2).contiguous().requires_grad_(True)
        v = v2.view(B, S, Hkv, D).transpose(1, 2).contiguous().requires_grad_(True)
        do = dout2.view(B, S, Hq, D).transpose(1, 2).contiguous()
        mask = None
        if P > 0:
            idx = torch.arange(S, device=dev)
            mask = (idx[None, :] < P) | (idx[:, None] >= idx[None, :])
        if "attn" in which:
            for name, backend in (("SDPA cuDNN", SDPBackend.CUDNN_ATTENTION), ("SDPA flash", SDPBackend.FLASH_ATTENTION),
                                  ("SDPA mem-efficient", SDPBackend.EFFICIENT_ATTENTION)):
                try:
                    with sdpa_kernel(backend):
                        def fwd():
                            if mask is None:
                                return F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
                            return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, enable_gqa=True)
                        out = fwd()
                        t_f = timeit(fwd)

                        def fb():
                            out_ = fwd()
                            out_.backward(do)
                            q.grad = k.grad = v.grad = None
                        t_fb = timeit(fb)
                    t_b = t_fb - t_f
                    print(f"B={B} S={S} P={P}  {name:20s} fwd {t_f*1e3:8.1f} us {fl/t_f/1e9:7.1f} TF/s | bwd {t_b*1e3:8.1f} us {2.5*fl/t_b/1e9:7.1f} TF/s  (bwd = fwd+bwd - fwd){'  dense bool mask' if mask is not None else ''}", flush=True)
                except Exception as e:
                    print(f"B={B} S={S} P={P}  {name:20s} unavailable: {str(e)[:100]}", flush=True)
        if "flex" in which:
            try:
                from torch.nn.attention.flex_attention import create_block_mask, flex_attention

                cflex = torch.compile(flex_attention)
                Pc = P

                def mask_mod(b, h, q_idx, kv_idx):
                    return (kv_idx < Pc) | (q_idx >= kv_idx)

                bm = create_block_mask(mask_mod, None, None, S, S, device=dev)
                rep = Hq // Hkv

                def ffwd():   # reference llama.py:129-132: K/V expanded to 32 heads first
                    return cflex(q, k.repeat_interleave(rep, 1), v.repeat_interleave(rep, 1), block_mask=bm)

                def ffwd_gqa():
                    return cflex(q, k, v, block_mask=bm, enable_gqa=True)

                for name, f in (("flex compiled (ref: repeat_interleave)", ffwd), ("flex compiled enable_gqa", ffwd_gqa)):
                    t0 = time.time()
                    f()
                    torch.cuda.synchronize()
                    tc = time.time() - t0
                    t_f = timeit(f)

                    def fb():
                        f().backward(do)
                        q.grad = k.grad = v.grad = None
                    fb()
                    t_fb = timeit(fb)
                    t_b = t_fb - t_f
                    print(f"B={B} S={S} P={P}  {name:38s} fwd {t_f*1e3:8.1f} us {fl/t_f/1e9:7.1f} TF/s | bwd {t_b*1e3:8.1f} us {2.5*fl/t_b/1e9:7.1f} TF/s  (first call {tc:.0f} s)", flush=True)
            except Exception as e:
                print(f"B={B} S={S} P={P}  flex_attention unavailable: {str(e)[:200]}", flush=True)
        del g, q, k, v, do, dqkv, dout2

if "gemm" in which:
    print("\n== INT8 row-scaled GEMM, M = 16384 (+ 4096): TOP/s = 2*M*N*K / time; all three produce the dequantised bf16 output "
          "except torch._int_mm (raw int32)")
    tri = None
    try:
        from baseline import ref_loader

        ref_loader.load(cpu_shim=False)
        from subclasses.int8_mm import int8_mm_dequant as tri
    except Exception as e:
        print("reference Triton kernel unavailable:", str(e)[:200])
    for (N, K) in [(4096, 4096), (14336, 4096), (4096, 14336)]:
        w8 = torch.randint(-127, 128, (N, K), device=dev, dtype=torch.int8)
        ws = torch.rand(N, device=dev).bfloat16()
        for M in (4096, 16384):
            xq = torch.randint(-127, 128, (M, K), device=dev, dtype=torch.int8)
            xs = torch.rand(M, device=dev).bfloat16()
            ops_ = 2.0 * M * N * K
            t_o = timeit(lambda: ops.int8_gemm_dequant(xq, w8, xs, ws))
            t_c = timeit(lambda: torch._int_mm(xq, w8.t()))
            line = f"{N:>6d} x {K:<6d} M={M:<6d} llamax_b200 {ops_/t_o/1e9:7.0f} T/s | cuBLASLt _int_mm {ops_/t_c/1e9:7.0f} T/s"
            if tri is not None:
                try:
                    t0 = time.time()
                    ref = tri(xq, w8.t(), xs, ws)
                    torch.cuda.synchronize()
                    tc = time.time() - t0
                    t_t = timeit(lambda: tri(xq, w8.t(), xs, ws))
                    same = torch.equal(ref, ops.int8_gemm_dequant(xq, w8, xs, ws))
                    line += f" | reference Triton kernel {ops_/t_t/1e9:7.0f} T/s (autotune {tc:.0f} s; output bit-identical to ours: {same})"
                except Exception as e:
                    line += f" | reference Triton kernel failed: {str(e)[:120]}"
            print(line, flush=True)
            del xq
        del w8

if "peak" in which:
    n = 8192
    a8 = torch.randint(-127, 128, (n, n), device=dev, dtype=torch.int8)
    b8 = torch.randint(-127, 128, (n, n), device=dev, dtype=torch.int8).t()
    a16 = torch.randn(n, n, device=dev).bfloat16()
    b16 = torch.randn(n, n, device=dev).bfloat16()
    ops_ = 2.0 * n ** 3
    t8 = timeit(lambda: torch._int_mm(a8, b8), reps=1, n=10)
    t16 = timeit(lambda: torch.matmul(a16, b16), reps=1, n=10)
    t8s = timeit(lambda: torch._int_mm(a8, b8), reps=3000, n=1)
    t16s = timeit(lambda: torch.matmul(a16, b16), reps=1500, n=1)
    xs = torch.rand(n, device=dev).bfloat16()
    to = timeit(lambda: ops.int8_gemm_dequant(a8, b8.t(), xs, xs), reps=1, n=10)
    tos = timeit(lambda: ops.int8_gemm_dequant(a8, b8.t(), xs, xs), reps=3000, n=1)
    tb = timeit(lambda: ops.bf16_gemm(a16, b16), reps=1, n=10)
    tbs = timeit(lambda: ops.bf16_gemm(a16, b16), reps=1500, n=1)
    print(f"\n== 8192^3 peaks: cuBLASLt int8 burst {ops_/t8/1e9:.0f} / sustained {ops_/t8s/1e9:.0f} TOP/s; cuBLAS bf16 burst {ops_/t16/1e9:.0f} / sustained {ops_/t16s/1e9:.0f} TFLOP/s")
    print(f"   llamax_b200: int8+dequant burst {ops_/to/1e9:.0f} / sustained {ops_/tos/1e9:.0f} TOP/s; bf16 burst {ops_/tb/1e9:.0f} / sustained {ops_/tbs/1e9:.0f} TFLOP/s")
print("ok")
