"""GPU parity of the HBM-bound passes (row quantisation bit-exact; norm / SwiGLU / RoPE against the reference
formulas; backward passes within 1e-2 of the fp32 evaluation)."""
import pytest
import torch

from llamax_b200 import ops
from oracle import ref_ops as R
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


# M >= 1024 with K <= 4096 runs the warp-per-row kernel (K = 32 * 8 * {1, 2, 4, 8, 16}: the branch-free kFull form; 3000, 1792:
# the bounds-checked one), the others the block-per-row / ring kernels
@pytest.mark.parametrize("M,K", [(64, 512), (300, 4096), (128, 14336), (5, 1792), (1, 8), (3, 32768), (1024, 4096),
                                 (2048, 512), (257, 1792), (300, 2056), (4100, 8), (1100, 3000), (1030, 1792),
                                 (1500, 2048)])
def test_rowquant_bit_exact(M, K):
    g = torch.Generator().manual_seed(K)
    x = (torch.randn(M, K, generator=g) * 3).bfloat16()
    x[0].zero_()                                   # all-zero row -> scale 0, codes 0
    if M > 2:
        x[1] = 1e-30                               # below the 1e-12 clip
        x[2, 0] = 3.0e38                           # huge dynamic range
    q_ref, s_ref = R.quantize_int8_rowwise(x)
    q, s = ops.rowquant_int8(x.cuda())
    assert torch.equal(q.cpu(), q_ref) and torch.equal(s.cpu(), s_ref)


def test_rowquant_exact_ties_round_half_even():
    """Quotients that sit exactly on half-integers (and one bf16 ulp either side): row maxima 127, 254 and 63.5 give
    scales 1, 2 and 0.5, so x / scale = k + 0.5 exactly; torch.round is half-to-even. Also a row whose scale is not a
    power of two, filled with the bf16 neighbours of every half-integer multiple of it."""
    k = torch.arange(-127, 127, dtype=torch.float32) + 0.5            # 254 exact ties, all bf16-representable
    rows = []
    for s in (1.0, 2.0, 0.5):
        r = torch.zeros(512)
        r[:254] = k * s
        r[254] = 127.0 * s                                            # pins the scale
        rows.append(r)
    s = 3.0 / 127.0
    near = (k * s).bfloat16().float()
    ulp = torch.maximum(near.abs(), torch.tensor(1e-3)) * 2.0 ** -8
    r = torch.zeros(512)
    r[:254] = near
    r[254:508] = (near + ulp).bfloat16().float()
    r[508] = 3.0
    rows.append(r)
    x = torch.stack(rows).bfloat16()
    q_ref, s_ref = R.quantize_int8_rowwise(x)
    assert q_ref[0, 0] == -126 and q_ref[0, 1] == -126 and q_ref[0, 127] == 0 and q_ref[0, 128] == 2   # half-to-even
    q, sc = ops.rowquant_int8(x.cuda())
    assert torch.equal(q.cpu(), q_ref) and torch.equal(sc.cpu(), s_ref)
    xb = x.repeat(128, 1)                                              # 512 rows: the warp-per-row kernel
    q, sc = ops.rowquant_int8(xb.cuda())
    assert torch.equal(q.cpu(), q_ref.repeat(128, 1)) and torch.equal(sc.cpu(), s_ref.repeat(128))


def test_rowquant_pitched_input():
    x = torch.randn(40, 1024).bfloat16()
    q, s = ops.rowquant_int8(x.cuda()[:, 256:768])
    q_ref, s_ref = R.quantize_int8_rowwise(x[:, 256:768].contiguous())
    assert torch.equal(q.cpu(), q_ref) and torch.equal(s.cpu(), s_ref)


# M >= 1024, D <= 4096: warp-per-row forward; M >= 592, D <= 4096: ring backward; the others the block-per-row kernels
@pytest.mark.parametrize("M,D", [(64, 512), (300, 4096), (7, 8192), (1024, 4096), (2048, 512), (300, 2056), (1100, 3000),
                                 (700, 4096)])
def test_rmsnorm_fwd_bwd(M, D):
    g = torch.Generator().manual_seed(D)
    x = torch.randn(M, D, generator=g).bfloat16()
    w = (1 + 0.1 * torch.randn(D, generator=g)).bfloat16()
    y_ref = R.rmsnorm_ref(x, w)
    y, rstd, q8, qs = ops.rmsnorm_fwd(x.cuda(), w.cuda(), 1e-5, quant=True)
    assert (y.cpu() != y_ref).float().mean().item() < 1e-3       # 1-ulp flips from the reduction order only
    assert rel_err(y, y_ref.float()) < 8e-3
    q_ref, s_ref = R.quantize_int8_rowwise(y.cpu())               # fused quantisation == quantising y
    assert torch.equal(q8.cpu(), q_ref) and torch.equal(qs.cpu(), s_ref)
    dy = torch.randn(M, D, generator=g).bfloat16()
    dres = torch.randn(M, D, generator=g).bfloat16()
    dx_ref, dw_ref = R.rmsnorm_bwd_f32(dy, x, w)
    dx, dw = ops.rmsnorm_bwd(dy.cuda(), x.cuda(), w.cuda(), rstd, dres.cuda())
    assert rel_err(dx, dx_ref + dres.float()) <= 1e-2 and rel_err(dw, dw_ref) <= 1e-2
    dx2, dw2 = ops.rmsnorm_bwd(dy.cuda(), x.cuda(), w.cuda(), rstd, None, want_dw=False)
    assert dw2 is None and rel_err(dx2, dx_ref) <= 1e-2


# (600, 14336), (2400, 1792), (2400, 1000): the persistent ring kernel (enough rows per resident CTA); the others one CTA per row
@pytest.mark.parametrize("M,F", [(64, 1792), (100, 14336), (600, 14336), (2400, 1792), (2400, 1000)])
def test_swiglu_fwd_bwd(M, F):
    g = torch.Generator().manual_seed(F)
    ab = (torch.randn(M, 2 * F, generator=g) * 2).bfloat16()
    a, b = ab[:, :F], ab[:, F:]
    g_ref = R.swiglu_ref(a, b)
    abc = ab.cuda()
    gg, q8, qs = ops.swiglu_fwd(abc[:, :F], abc[:, F:], quant=True)
    assert (gg.cpu() != g_ref).float().mean().item() < 1e-3
    q_ref, s_ref = R.quantize_int8_rowwise(gg.cpu())
    assert torch.equal(q8.cpu(), q_ref) and torch.equal(qs.cpu(), s_ref)
    none_g, q8b, qsb = ops.swiglu_fwd(abc[:, :F], abc[:, F:], quant=True, want_g=False)      # codes without writing g
    assert none_g is None and torch.equal(q8b, q8) and torch.equal(qsb, qs)
    g_only, q_none, _ = ops.swiglu_fwd(abc[:, :F], abc[:, F:], quant=False)                     # g without the quantiser
    assert q_none is None and torch.equal(g_only, gg)
    a_sep, b_sep = abc[:, :F].contiguous(), abc[:, F:].contiguous()                            # two allocations, pitch F
    g_sep, q_sep, s_sep = ops.swiglu_fwd(a_sep, b_sep, quant=True)
    assert torch.equal(g_sep, gg) and torch.equal(q_sep, q8) and torch.equal(s_sep, qs)
    dg = torch.randn(M, F, generator=g).bfloat16()
    da_ref, db_ref = R.swiglu_bwd_f32(dg, a, b)
    wide = torch.zeros(M, 2 * F + 16, dtype=torch.bfloat16, device="cuda")
    da, db, g2 = ops.swiglu_bwd(dg.cuda(), abc[:, :F], abc[:, F:], want_g=True, out_ab=wide)
    assert rel_err(da, da_ref) <= 1e-2 and rel_err(db, db_ref) <= 1e-2
    assert torch.equal(g2.cpu(), gg.cpu()) and (wide[:, 2 * F :] == 0).all()


def test_rope_bit_exact_and_inverse():
    B, S, H, D = 2, 300, 6, 128
    rope = R.build_rope(D, 512, 500000, True)
    x = torch.randn(B, S, H, D).bfloat16()
    y_ref = R.apply_rope(x, rope[:S])
    xc = torch.cat([x.reshape(B * S, H * D), torch.zeros(B * S, 256).bfloat16()], 1).cuda()
    ops.rope_(xc, rope.cuda(), B, S, H, D)
    assert torch.equal(xc[:, : H * D].cpu().view(B, S, H, D), y_ref)
    assert (xc[:, H * D :] == 0).all()
    dy = torch.randn(B, S, H, D).bfloat16()
    dyc = dy.reshape(B * S, H * D).clone().cuda()
    ops.rope_(dyc, rope.cuda(), B, S, H, D, inverse=True)
    assert torch.equal(dyc.cpu().view(B, S, H, D), R.apply_rope_inverse(dy, rope[:S]))


def test_dequant_weight_and_lora_wgrad():
    N, K = 264, 256
    w8 = torch.randint(-127, 128, (N, K), dtype=torch.int8)
    s = (torch.rand(N) * 0.01).bfloat16()
    o1 = ops.dequant_weight(w8.cuda(), s.cuda(), transpose=False, apply_scale=False).cpu()
    assert torch.equal(o1, w8.bfloat16())
    wide = torch.zeros(K, N + 24, dtype=torch.bfloat16, device="cuda")
    ops.dequant_weight(w8.cuda(), s.cuda(), transpose=True, apply_scale=True, out=wide[:, :N])
    assert torch.equal(wide[:, :N].cpu(), (w8.float() * s.float()[:, None]).bfloat16().T) and (wide[:, N:] == 0).all()
    for Rr in (8, 16, 24):
        M, Pn = 1000, 512
        X, Hh = torch.randn(M, Pn).bfloat16(), torch.randn(M, Rr).bfloat16()
        o = ops.lora_wgrad(X.cuda(), Hh.cuda(), 0.5).cpu()
        assert rel_err(o, 0.5 * X.float().T @ Hh.float()) < 1e-4


@pytest.mark.parametrize("M,V", [(37, 1024), (16, 128256)])
def test_cross_entropy_fwd_bwd(M, V):
    """F.cross_entropy(logits.float(), labels) with ignore_index -100 (llama.py:216-218), loss and d/dlogits."""
    g = torch.Generator().manual_seed(V)
    logits = (torch.randn(M, V, generator=g) * 3).bfloat16()
    labels = torch.randint(0, V, (M,), generator=g)
    labels[::5] = -100
    lf = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf, labels)
    ref.backward()
    lc = logits.cuda().clone()
    loss_sum = torch.zeros((), device="cuda", dtype=torch.float32)
    inv_n = torch.tensor(1.0 / (labels != -100).sum().item(), device="cuda", dtype=torch.float32)
    ops.cross_entropy_(lc, labels.cuda(), loss_sum, inv_n, True)
    assert abs((loss_sum * inv_n).item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert rel_err(lc, lf.grad) <= 1e-2
    assert (lc[::5] == 0).all()


def test_batched_copy_jobs():
    """llamax_batched_copy: mixed bf16 / fp32 sources, transposes, scales, pitched sources and destinations, ragged
    tile edges — bit-exact against bf16(scale * src) (same single rounding)."""
    torch.manual_seed(3)
    dev = "cuda"
    a = torch.randn(8, 4096, device=dev).bfloat16()                       # A [R,K] -> A^T into a pitched operand slice
    wt = torch.zeros(4096, 6168, device=dev, dtype=torch.bfloat16)
    b = torch.randn(1000, 16, device=dev).bfloat16()                      # B [N,R] -> scale * B^T
    bt = torch.empty(16, 1000, device=dev, dtype=torch.bfloat16)
    h = torch.randn(300, 24, device=dev).bfloat16()[:, 8:16]              # strided h slice -> h^T with padded pitch
    ht = ops.transposed_rank_buffer(8, 300, dev)
    g = torch.randn(4096, 24, device=dev)                                 # fp32 dA^T slice -> bf16 dA
    da = torch.empty(8, 4096, device=dev, dtype=torch.bfloat16)
    f = torch.randn(333, 8, device=dev)                                   # fp32 dB -> bf16 dB
    db = torch.empty(333, 8, device=dev, dtype=torch.bfloat16)
    ops.batched_copy([(a, wt[:, 6144:6152], 1.0, True), (b, bt, 0.5, True), (h, ht, 1.0, True),
                      (g[:, 16:24], da, 1.0, True), (f, db, 1.0, False)])
    assert torch.equal(wt[:, 6144:6152], a.t()) and torch.count_nonzero(wt[:, :6144]) == 0
    assert torch.equal(bt, (b.float() * 0.5).bfloat16().t())
    assert torch.equal(ht, h.t())
    assert torch.equal(da, g[:, 16:24].bfloat16().t())
    assert torch.equal(db, f.bfloat16())
    jobs = [(b, torch.empty(16, 1000, device=dev, dtype=torch.bfloat16), 1.0, True) for _ in range(70)]  # > 64 jobs
    ops.batched_copy(jobs)
    assert all(torch.equal(j[1], b.t()) for j in jobs)
