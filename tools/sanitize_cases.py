"""One small launch of every kernel of the library (the command run under compute-sanitizer, one tool per gpurun call:
memcheck / racecheck / synccheck; logs under profiles/). Sizes are the smallest that still exercise every code path:
ragged tails, both CTA-group sizes, epilogue variants, masked / unmasked attention tiles, packed documents."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from llamax_b200 import ops

torch.manual_seed(0)
dev = "cuda"
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()

# ---- GEMMs
for cg in (2, 1):
    ops.set_gemm_cta_group(cg)
    M, N, K = 300, 264, 208
    A8 = torch.randint(-127, 128, (M, K), device=dev, dtype=torch.int8)
    W8 = torch.randint(-127, 128, (N, K), device=dev, dtype=torch.int8)
    sa, sw = torch.rand(M, device=dev).bfloat16(), torch.rand(N, device=dev).bfloat16()
    ops.int8_gemm_s32(A8, W8)
    ops.int8_gemm_dequant(A8, W8, sa, sw)
    ops.int8_gemm_dequant(A8, W8, sa, sw, lora_h=bf(M, 8), lora_b=bf(N, 8), lora_scale=0.5, resid=bf(M, N))
    ops.bf16_gemm(bf(M, K), bf(N, K))
    ops.bf16_gemm(bf(M, K), bf(N, K), col_scale=sw, round_before_scale=True, lora_h=bf(M, 16), lora_b=bf(N, 16), resid=bf(M, N))
    ops.bf16_gemm(bf(M, K), bf(8, K))                      # skinny
    ops.bf16_gemm_tn(bf(K, 256), bf(K, 264))
    F_ = 272
    ab = bf(M, 2 * F_)
    ops.bf16_gemm_swiglu_bwd(bf(M, K), bf(F_, K), ab[:, :F_], ab[:, F_:], out_ab=torch.empty(M, 2 * F_ + 16, device=dev, dtype=torch.bfloat16),
                             want_g=True, lora_h=bf(M, 8), lora_b=bf(F_, 8))
ops.set_gemm_cta_group(2)
ops.bf16_gemm(bf(1100, 8200), bf(264, 8200))               # wide 512 x 256 kernel
ops.dequant_weight(torch.randint(-127, 128, (264, 208), device=dev, dtype=torch.int8), torch.rand(264, device=dev).bfloat16(),
                   transpose=True, apply_scale=True)

# ---- LoRA kernels
M, N, R = 301, 520, 8
ops.lora_wgrad(bf(M, N), bf(M, R), 0.5)
ops.lora_bwd_pair(bf(M, N), bf(R, N), bf(M, R), torch.empty(M, R, device=dev, dtype=torch.bfloat16), 1.0)

# ---- elementwise
M, D, F_ = 67, 512, 1792
x, w = bf(M, D), bf(D)
y, rstd, q8, qs = ops.rmsnorm_fwd(x, w, 1e-5, quant=True)
ops.rmsnorm_bwd(bf(M, D), x, w, rstd, bf(M, D))
M2 = 4096                                                    # enough rows for the persistent "ring" variants
x2 = bf(M2, D)
y2, rstd2, _, _ = ops.rmsnorm_fwd(x2, w, 1e-5, quant=True)
ops.rowquant_int8(x2)
ops.rowquant_int8_colscale(x, w)
ab = bf(M, 2 * F_)
ops.swiglu_fwd(ab[:, :F_], ab[:, F_:], quant=True, want_g=True)
ops.swiglu_bwd(bf(M, F_), ab[:, :F_], ab[:, F_:], want_g=True)
B, S, H, Dh = 2, 40, 4, 128
rope = torch.randn(64, Dh // 2, 2, device=dev)
ops.rope_(bf(B * S, H * Dh), rope, B, S, H, Dh)
logits = bf(37, 1024)
ops.cross_entropy_(logits, torch.randint(0, 1024, (37,), device=dev), torch.zeros((), device=dev), torch.ones((), device=dev), True)
z = bf(2 * 34, 64)
ops.gelu_bias_fwd_(z, bf(64), 34, 1, 33)
ops.gelu_bwd(bf(2 * 34, 64), z, 34, 1, 33)
ops.conv_s2k3_col2im(bf(2 * 17, 3 * 64), 2, 34, 64)
ops.batched_copy([(bf(8, 40), torch.empty(40, 8, device=dev, dtype=torch.bfloat16), 0.5, True)])

# ---- attention: causal, prefix-LM (ragged S), packed documents; head_dim 128 and 64
for (B, S, Hq, Hkv, Dh, P, docs) in ((1, 300, 4, 2, 128, 0, False), (2, 301, 2, 1, 128, 100, False), (1, 384, 2, 1, 128, 0, True),
                                     (1, 256, 2, 2, 64, 64, False)):
    ld = (Hq + 2 * Hkv) * Dh
    g = bf(B * S, ld)
    q, k, v = g[:, : Hq * Dh], g[:, Hq * Dh : (Hq + Hkv) * Dh], g[:, (Hq + Hkv) * Dh :]
    ds = de = None
    if docs:
        ids = torch.repeat_interleave(torch.arange(3, device=dev), torch.tensor([100, 156, 128], device=dev))[None]
        ds, de = ops.doc_bounds(ids)
    o, lse = ops.attn_fwd(q, k, v, B, S, Hq, Hkv, Dh, P, doc_start=ds)
    dqkv = torch.empty_like(g)
    ops.attn_bwd(q, k, v, o, lse, bf(B * S, Hq * Dh), dqkv[:, : Hq * Dh], dqkv[:, Hq * Dh : (Hq + Hkv) * Dh],
                 dqkv[:, (Hq + Hkv) * Dh :], B, S, Hq, Hkv, Dh, P, doc_start=ds, doc_end=de, rope_inverse=None)
for v_ in ("2", "1"):   # the other forward variants are selected per process; exercised by the A/B switch in their own runs
    pass
torch.cuda.synchronize()
print("sanitize_cases: ok")
