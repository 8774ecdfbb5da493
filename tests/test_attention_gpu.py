"""GPU parity of the prefix-LM attention kernels (forward + backward) against the dense fp64 oracle, plus
size-independent properties at BASELINE config-5 sizes."""
import pytest
import torch

from llamax_b200 import ops
from oracle import ref_ops as R
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
D = 128


def _split(qkv, Hq, Hkv):
    return qkv[:, : Hq * D], qkv[:, Hq * D : (Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D :]


CASES = [(1, 128, 1, 1, 0), (1, 256, 4, 1, 0), (2, 512, 8, 2, 128), (1, 300, 4, 2, 70), (1, 1628, 4, 1, 1500),
         (1, 2048, 8, 2, 0), (1, 640, 4, 4, 640), (3, 129, 2, 1, 1), (1, 1024, 4, 1, 768)]


@pytest.mark.parametrize("B,S,Hq,Hkv,P", CASES)
def test_attention_fwd_bwd_vs_oracle(B, S, Hq, Hkv, P):
    torch.manual_seed(S + P)
    qkv = torch.randn(B * S, (Hq + 2 * Hkv) * D).bfloat16()
    dout = torch.randn(B * S, Hq * D).bfloat16()
    q, k, v = _split(qkv, Hq, Hkv)
    to4 = lambda t, H: t.reshape(B, S, H, D).transpose(1, 2)
    o_ref, dq_ref, dk_ref, dv_ref = R.attention_ref_grads(to4(q, Hq), to4(k, Hkv), to4(v, Hkv), to4(dout, Hq), P)
    g = qkv.cuda()
    qc, kc, vc = _split(g, Hq, Hkv)
    o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, P)
    assert rel_err(to4(o.cpu(), Hq), o_ref) <= 1e-2
    s = (to4(q, Hq).double() @ to4(k, Hkv).double().repeat_interleave(Hq // Hkv, 1).transpose(-1, -2)) / D ** 0.5
    s = s.masked_fill(~R.prefix_lm_mask(S, P), float("-inf"))
    assert (lse.cpu().double() - torch.logsumexp(s, -1)).abs().max().item() < 1e-3
    dqkv = torch.zeros_like(g)
    dq, dk, dv = _split(dqkv, Hq, Hkv)
    ops.attn_bwd(qc, kc, vc, o, lse, dout.cuda(), dq, dk, dv, B, S, Hq, Hkv, D, P)
    assert rel_err(to4(dq.cpu(), Hq), dq_ref) <= 1e-2
    assert rel_err(to4(dk.cpu(), Hkv), dk_ref) <= 1e-2
    assert rel_err(to4(dv.cpu(), Hkv), dv_ref) <= 1e-2


@pytest.mark.parametrize("S,P", [(4096, 0), (4096, 1024), (16384, 0), (1756, 1500)])
def test_attention_properties_full_size(S, P):
    """Config-5 sizes (B*S = 16384, Hq=32, Hkv=8): (1) V = const -> O = const (softmax rows sum to 1);
    (2) masking: perturbing K/V at suffix positions >= t leaves O[:t] unchanged, and prefix rows (q < P) do not see
    the suffix at all; (3) dV column sums equal dO column sums when V = const... checked via linearity in V."""
    B, Hq, Hkv = max(1, 16384 // S) if S <= 4096 else 1, 32, 8
    B = min(B, 2)
    torch.manual_seed(S)
    g = torch.randn(B * S, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
    qc, kc, vc = _split(g, Hq, Hkv)
    o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, P)
    g1 = g.clone()
    _, _, v1 = _split(g1, Hq, Hkv)
    v1.fill_(0.5)
    o1, _ = ops.attn_fwd(*_split(g1, Hq, Hkv), B, S, Hq, Hkv, D, P)
    assert (o1.float() - 0.5).abs().max().item() < 4e-3
    t = max(P, S // 2) + 3
    g2 = g.clone()
    g2.view(B, S, -1)[:, t:, Hq * D :] = torch.randn(B, S - t, 2 * Hkv * D, device="cuda").bfloat16()  # K, V suffix
    o2, lse2 = ops.attn_fwd(*_split(g2, Hq, Hkv), B, S, Hq, Hkv, D, P)
    assert torch.equal(o2.view(B, S, -1)[:, :t], o.view(B, S, -1)[:, :t])
    assert torch.equal(lse2[:, :, :t], lse[:, :, :t])
    # linearity in V: O(V_a + V_b) = O(V_a) + O(V_b) (same P matrix), to bf16 rounding
    g3 = g.clone()
    _, _, v3 = _split(g3, Hq, Hkv)
    v3.mul_(2.0)
    o3, _ = ops.attn_fwd(*_split(g3, Hq, Hkv), B, S, Hq, Hkv, D, P)
    assert rel_err(o3, 2.0 * o.double()) < 1e-2


@pytest.mark.parametrize("B,S,Hq,Hkv,lengths", [
    (1, 640, 4, 1, [100, 28, 300, 212]),          # documents crossing 128-tile borders, one exactly tile-aligned end
    (1, 2048, 8, 2, [700, 1, 600, 747]),          # a 1-token document
    (2, 384, 2, 1, [384]),                        # single document == plain causal
    (1, 1000, 4, 2, [128, 128, 256, 488]),        # tile-aligned documents, ragged total
])
def test_document_causal_mask_fwd_bwd(B, S, Hq, Hkv, lengths):
    """Packed-sequence document-causal mask of the reference trainer (train_metamathqa.py:51-83): fully masked tiles
    are skipped (range from doc_start / doc_end), partial tiles use the per-row document start."""
    assert sum(lengths) == S
    doc_ids = torch.repeat_interleave(torch.arange(len(lengths)), torch.tensor(lengths))[None].expand(B, S).contiguous()
    torch.manual_seed(S)
    qkv = torch.randn(B * S, (Hq + 2 * Hkv) * D).bfloat16()
    dout = torch.randn(B * S, Hq * D).bfloat16()
    q, k, v = _split(qkv, Hq, Hkv)
    to4 = lambda t, H: t.reshape(B, S, H, D).transpose(1, 2)
    o_ref, dq_ref, dk_ref, dv_ref = R.attention_ref_grads(to4(q, Hq), to4(k, Hkv), to4(v, Hkv), to4(dout, Hq), 0,
                                                          doc_ids=doc_ids)
    ds, de = ops.doc_bounds(doc_ids.cuda())
    starts = torch.tensor([0] + lengths[:-1]).cumsum(0)
    assert torch.equal(ds[0].cpu(), torch.repeat_interleave(starts, torch.tensor(lengths)).int())
    assert torch.equal(de[0].cpu(), torch.repeat_interleave(starts + torch.tensor(lengths) - 1, torch.tensor(lengths)).int())
    g = qkv.cuda()
    qc, kc, vc = _split(g, Hq, Hkv)
    o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, 0, doc_start=ds)
    assert rel_err(to4(o.cpu(), Hq), o_ref) <= 1e-2
    dqkv = torch.zeros_like(g)
    dq, dk, dv = _split(dqkv, Hq, Hkv)
    ops.attn_bwd(qc, kc, vc, o, lse, dout.cuda(), dq, dk, dv, B, S, Hq, Hkv, D, 0, doc_start=ds, doc_end=de)
    assert rel_err(to4(dq.cpu(), Hq), dq_ref) <= 1e-2
    assert rel_err(to4(dk.cpu(), Hkv), dk_ref) <= 1e-2
    assert rel_err(to4(dv.cpu(), Hkv), dv_ref) <= 1e-2
    if len(lengths) == 1:  # identical to the plain causal path, bit for bit
        o2, lse2 = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, 0)
        assert torch.equal(o, o2) and torch.equal(lse, lse2)


@pytest.mark.parametrize("B,S,Hq,Hkv,P", [(2, 1024, 8, 2, 512), (1, 300, 8, 2, 0), (1, 129, 2, 1, 77)])
def test_attention_head_dim_64(B, S, Hq, Hkv, P):
    """head_dim 64 (BASELINE config 0: dim 512, GQA 8/2): same kernels on TMA-zero-padded 128-wide tiles."""
    Dh = 64
    torch.manual_seed(S)
    qkv = torch.randn(B * S, (Hq + 2 * Hkv) * Dh).bfloat16()
    dout = torch.randn(B * S, Hq * Dh).bfloat16()
    sp = lambda t: (t[:, : Hq * Dh], t[:, Hq * Dh : (Hq + Hkv) * Dh], t[:, (Hq + Hkv) * Dh :])
    q, k, v = sp(qkv)
    to4 = lambda t, H: t.reshape(B, S, H, Dh).transpose(1, 2)
    o_ref, dq_ref, dk_ref, dv_ref = R.attention_ref_grads(to4(q, Hq), to4(k, Hkv), to4(v, Hkv), to4(dout, Hq), P)
    g = qkv.cuda()
    qc, kc, vc = sp(g)
    o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, Dh, P)
    assert rel_err(to4(o.cpu(), Hq), o_ref) <= 1e-2
    dqkv = torch.zeros_like(g)
    dq, dk, dv = sp(dqkv)
    ops.attn_bwd(qc, kc, vc, o, lse, dout.cuda(), dq, dk, dv, B, S, Hq, Hkv, Dh, P)
    assert rel_err(to4(dq.cpu(), Hq), dq_ref) <= 1e-2
    assert rel_err(to4(dk.cpu(), Hkv), dk_ref) <= 1e-2
    assert rel_err(to4(dv.cpu(), Hkv), dv_ref) <= 1e-2


@pytest.mark.parametrize("B,S,Hq,Hkv,P,Dh", [(2, 384, 8, 2, 100, 128), (1, 300, 4, 2, 0, 64)])
def test_attention_bwd_fused_rope_backward(B, S, Hq, Hkv, P, Dh):
    """attn_bwd(rope_inverse=table) == attn_bwd followed by the in-place inverse RoPE pass on the q|k gradients
    (autograd of apply_rope, llama.py:63-73). The fused path rotates the fp32 sums, so it differs from the two-pass
    result only by bf16 rounding; both are checked against the fp64 rotation of the oracle gradients."""
    torch.manual_seed(5 + S)
    nq, nk = Hq * Dh, Hkv * Dh
    qkv = torch.randn(B * S, nq + 2 * nk).bfloat16()
    dout = torch.randn(B * S, nq).bfloat16()
    rope = R.build_rope(Dh, 512, 500000.0, True)[:S].contiguous()
    to4 = lambda t, H: t.reshape(B, S, H, Dh).transpose(1, 2)
    _, dq_ref, dk_ref, dv_ref = R.attention_ref_grads(to4(qkv[:, :nq], Hq), to4(qkv[:, nq:nq + nk], Hkv),
                                                      to4(qkv[:, nq + nk:], Hkv), to4(dout, Hq), P)

    def unrotate(g):  # [B,H,S,D] fp64 -> RoPE backward
        g = g.double()
        c, s = rope[:, :, 0].double(), rope[:, :, 1].double()  # [S, D/2]
        g0, g1 = g[..., 0::2], g[..., 1::2]
        return torch.stack((g0 * c + g1 * s, g1 * c - g0 * s), -1).flatten(-2)

    g = qkv.cuda()
    qc, kc, vc = g[:, :nq], g[:, nq:nq + nk], g[:, nq + nk:]
    o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, Dh, P)
    fused = torch.zeros_like(g)
    ops.attn_bwd(qc, kc, vc, o, lse, dout.cuda(), fused[:, :nq], fused[:, nq:nq + nk], fused[:, nq + nk:], B, S, Hq, Hkv,
                 Dh, P, rope_inverse=rope.cuda())
    two = torch.zeros_like(g)
    ops.attn_bwd(qc, kc, vc, o, lse, dout.cuda(), two[:, :nq], two[:, nq:nq + nk], two[:, nq + nk:], B, S, Hq, Hkv, Dh, P)
    ops.rope_(two, rope.cuda(), B, S, Hq + Hkv, Dh, inverse=True)
    assert torch.equal(fused[:, nq + nk:], two[:, nq + nk:])                     # dv untouched
    assert rel_err(fused[:, :nq + nk].float().cpu(), two[:, :nq + nk].float().cpu()) <= 1e-2
    assert rel_err(to4(fused[:, :nq].cpu(), Hq), unrotate(dq_ref)) <= 1e-2
    assert rel_err(to4(fused[:, nq:nq + nk].cpu(), Hkv), unrotate(dk_ref)) <= 1e-2


@pytest.mark.parametrize("B,S,Hq,Hkv,P", [(2, 301, 4, 2, 100), (1, 128, 2, 1, 0)])
def test_attention_backward_writes_stay_inside_their_views(B, S, Hq, Hkv, P):
    """Canary test (compute-sanitizer is closed on this pool): dq / dk / dv are column views of a larger buffer filled
    with a sentinel, with spare rows after the last token; nothing outside the three views may change, ragged S
    included (the kernels' row / column guards)."""
    torch.manual_seed(3)
    ld = (Hq + 2 * Hkv) * D
    g = torch.randn(B * S, ld, device="cuda").bfloat16()
    q, k, v = _split(g, Hq, Hkv)
    o, lse = ops.attn_fwd(q, k, v, B, S, Hq, Hkv, D, P)
    pad = 16
    big = torch.full((B * S + 8, ld + 2 * pad), 3.0, device="cuda", dtype=torch.bfloat16)
    inner = big[: B * S, pad : pad + ld]
    dq, dk, dv = _split(inner, Hq, Hkv)
    ops.attn_bwd(q, k, v, o, lse, torch.randn(B * S, Hq * D, device="cuda").bfloat16(), dq, dk, dv, B, S, Hq, Hkv, D, P)
    torch.cuda.synchronize()
    assert bool((big[B * S :] == 3).all()) and bool((big[:, :pad] == 3).all()) and bool((big[:, pad + ld :] == 3).all())
    assert not bool((inner == 3).all())
    # forward outputs: every row written, lse finite
    assert bool(torch.isfinite(lse).all()) and bool(torch.isfinite(o.float()).all())


def test_per_sequence_prefix_lengths_fwd_bwd():
    """One prefix length per sequence of the batch (int32 [B] through the C ABI): every sequence matches the oracle run
    with its own scalar prefix; a constant vector equals the scalar call bit for bit."""
    B, S, Hq, Hkv = 3, 300, 4, 2
    Ps = [0, 77, 300]
    torch.manual_seed(8)
    qkv = torch.randn(B * S, (Hq + 2 * Hkv) * D).bfloat16()
    dout = torch.randn(B * S, Hq * D).bfloat16()
    g = qkv.cuda()
    qc, kc, vc = _split(g, Hq, Hkv)
    pb = torch.tensor(Ps, device="cuda", dtype=torch.int32)
    o, lse = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, pb)
    dqkv = torch.zeros_like(g)
    dq, dk, dv = _split(dqkv, Hq, Hkv)
    ops.attn_bwd(qc, kc, vc, o, lse, dout.cuda(), dq, dk, dv, B, S, Hq, Hkv, D, pb)
    q, k, v = _split(qkv, Hq, Hkv)
    to4 = lambda t, H, b: t.reshape(B, S, H, D)[b : b + 1].transpose(1, 2)
    for b, P in enumerate(Ps):
        o_ref, dq_ref, dk_ref, dv_ref = R.attention_ref_grads(to4(q, Hq, b), to4(k, Hkv, b), to4(v, Hkv, b), to4(dout, Hq, b), P)
        assert rel_err(to4(o.cpu(), Hq, b), o_ref) <= 1e-2
        assert rel_err(to4(dq.cpu(), Hq, b), dq_ref) <= 1e-2
        assert rel_err(to4(dk.cpu(), Hkv, b), dk_ref) <= 1e-2
        assert rel_err(to4(dv.cpu(), Hkv, b), dv_ref) <= 1e-2
    o1, lse1 = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, 77)
    o2, lse2 = ops.attn_fwd(qc, kc, vc, B, S, Hq, Hkv, D, torch.full((B,), 77, device="cuda", dtype=torch.int32))
    assert torch.equal(o1, o2) and torch.equal(lse1, lse2)


def test_generic_flex_block_mask_is_recognised_by_the_fused_block():
    """A FlexAttention BlockMask built from a user mask_mod (no llamax tags) and a dense boolean mask drive the fused block
    exactly like PrefixLM(P); a sliding-window mask_mod raises."""
    from torch.nn.attention.flex_attention import create_block_mask

    from llamax_b200.modelling import PrefixLM
    from tests.helpers import build_tiny_llama

    P, B, S = 100, 2, 256
    model = build_tiny_llama(True, num_layers=1).cuda()
    layer, cfg = model.layers[0], model.config
    rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:S].cuda()
    x = torch.randn(B, S, cfg.embed_dim, device="cuda").bfloat16()
    want = layer(x, rope, block_mask=PrefixLM(P))
    bm = create_block_mask(lambda b, h, q, kv: (kv < P) | (q >= kv), None, None, S, S, device="cuda")
    assert torch.equal(layer(x, rope, block_mask=bm), want)
    idx = torch.arange(S, device="cuda")
    dense = ((idx[None, :] < P) | (idx[:, None] >= idx[None, :]))[None, None]
    assert torch.equal(layer(x, rope, mask=dense), want)
    sw = create_block_mask(lambda b, h, q, kv: (q >= kv) & (q - kv < 64), None, None, S, S, device="cuda")
    with pytest.raises(NotImplementedError):
        layer(x, rope, block_mask=sw)
