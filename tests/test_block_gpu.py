"""GPU parity: fused decoder block / tiny model (product path, C ABI) against the CPU oracle.

Tolerance (BASELINE.json north_star): max |ours - ref_fp32| / max |ref_fp32| <= 1e-2 for bf16 outputs and
gradients, where ref_fp32 is the reference op sequence evaluated in fp32; the bf16 reference itself is reported
beside it (it carries the same rounding noise as we do).
"""
import pytest
import torch

from oracle import ref_ops as R
from tests.helpers import TOL, build_tiny_llama, check_parity_table, oracle_layer_weights, rel_err

pytestmark = pytest.mark.gpu

# Documented exceptions to the strict 1e-2 bar: the PARAMETER gradients (LoRA A / B, RMSNorm weights) — sums over all
# tokens of products of bf16-rounded factors, on which the REFERENCE's own bf16 op sequence, evaluated on CPU on the same
# inputs, is itself 0.7 - 1.5e-2 away from its fp32 evaluation (tables in profiles/r2_parity_tables.txt). Block output and
# input gradient are NOT on the list: strict 1e-2. check_parity_table holds a listed tensor to max(1e-2, 1.5 x the
# reference's error in the same run).
KNOWN_BF16_LIMITED = ("a_wq", "a_wk", "a_wv", "a_wo", "a_w1", "a_w3", "a_w2", "an", "fn", "b_wq", "b_wk", "b_wv",
                      "b_wo", "b_w1", "b_w3", "b_w2")


def _layer_case(dynamic: bool, prefix_len: int, B=2, S=320, rank=8, adapters="all", config=None):
    model = build_tiny_llama(dynamic, num_layers=1, rank=rank, adapters=adapters, config=config)
    layer = model.layers[0]
    cfg = model.config
    rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:S]
    torch.manual_seed(3)
    x = torch.randn(B, S, cfg.embed_dim).bfloat16()
    dout = torch.randn(B, S, cfg.embed_dim).bfloat16()

    refs = {}
    for dtype in (torch.float32, torch.bfloat16):
        lw = oracle_layer_weights(layer, dtype)
        xr = x.detach().clone().to(dtype).requires_grad_(True)
        out = R.transformer_layer_ref(xr, rope, lw, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, prefix_len, dynamic)
        out.backward(dout.to(dtype))
        refs[dtype] = dict(out=out.detach(), dx=xr.grad, an=lw.attention_norm.grad, fn=lw.ffn_norm.grad,
                           **{f"a_{k}": v.grad for k, v in lw.lora_a.items()},
                           **{f"b_{k}": v.grad for k, v in lw.lora_b.items()})

    from llamax_b200.modelling import PrefixLM

    layer = layer.cuda()
    xc = x.detach().clone().cuda().requires_grad_(True)
    out = layer(xc, rope.cuda(), block_mask=PrefixLM(prefix_len))
    out.backward(dout.cuda())
    att, ff = layer.attention, layer.feed_forward
    mods = dict(wq=att.wq, wk=att.wk, wv=att.wv, wo=att.wo, w1=ff.w1, w3=ff.w3, w2=ff.w2)
    mods = {k: m for k, m in mods.items() if getattr(m, "rank", 0) > 0}
    ours = dict(out=out, dx=xc.grad, an=layer.attention_norm.weight.grad, fn=layer.ffn_norm.weight.grad,
                **{f"a_{k}": m.lora_a.grad for k, m in mods.items()},
                **{f"b_{k}": m.lora_b.grad for k, m in mods.items()})
    report = {}
    for key, val in ours.items():
        assert val is not None, key
        report[key] = (rel_err(val, refs[torch.float32][key]), rel_err(refs[torch.bfloat16][key], refs[torch.float32][key]))
    return report


@pytest.mark.parametrize("dynamic", [True, False])
@pytest.mark.parametrize("prefix_len", [0, 100])
def test_fused_block_matches_oracle(dynamic, prefix_len):
    report = _layer_case(dynamic, prefix_len)
    check_parity_table(report, f"fused block dim 512, dynamic={dynamic}, P={prefix_len}", KNOWN_BF16_LIMITED)


@pytest.mark.parametrize("dynamic", [True, False])
def test_fused_block_8b_shape_matches_oracle(dynamic):
    """ONE decoder block at the Llama-3.1-8B shape (D 4096, F 14336, 32 / 8 heads x 128, LoRA r = 8), 512 positions, prefix
    300, both INT8 modes, against the CPU oracle (fp32 evaluation of the reference op sequence): every output and
    gradient within the strict bar."""
    from llamax_b200.modelling import LlamaConfig

    cfg = LlamaConfig(4096, 1, 128, 32, 8, 14336, max_seq_len=512, vocab_size=1024, rope_base=500000, is_llama3_1=True)
    report = _layer_case(dynamic, 300, B=1, S=512, config=cfg)
    check_parity_table(report, f"fused block 8B shape, dynamic={dynamic}, P=300, S=512", KNOWN_BF16_LIMITED)


def test_int8_codes_at_8b_width_are_bit_exact():
    """The int8 activation codes entering the frozen projections at D = 4096 / F = 14336: codes and scales produced by the
    fused RMSNorm / SwiGLU passes equal quantize_int8_rowwise (int8.py:10-16) of the bf16 tensor the same pass writes,
    bit for bit, and that bf16 tensor differs from the oracle's by at most 1 bf16 ulp on < 0.1 % of the elements."""
    from llamax_b200 import ops

    torch.manual_seed(21)
    M, D, F_ = 512, 4096, 14336
    x = (torch.randn(M, D) * 1.7).bfloat16()
    w = (1 + 0.1 * torch.randn(D)).bfloat16()
    y, rstd, q8, qs = ops.rmsnorm_fwd(x.cuda(), w.cuda(), 1e-5, quant=True)
    q_ref, s_ref = R.quantize_int8_rowwise(y.cpu())
    assert torch.equal(q8.cpu(), q_ref) and torch.equal(qs.cpu(), s_ref.reshape(-1))
    y_ref = R.rmsnorm_ref(x, w)
    diff = (y.cpu().view(torch.int16).int() - y_ref.view(torch.int16).int()).abs()
    assert diff.max() <= 1 and (diff != 0).float().mean() < 1e-3
    ab = torch.randn(M, 2 * F_).bfloat16().cuda()
    g, gq, gs = ops.swiglu_fwd(ab[:, :F_], ab[:, F_:], quant=True, want_g=True)
    q_ref, s_ref = R.quantize_int8_rowwise(g.cpu())
    assert torch.equal(gq.cpu(), q_ref) and torch.equal(gs.cpu(), s_ref.reshape(-1))


def test_tiny_model_loss_and_grads():
    """4-layer tiny LlamaAudio-shaped text model: loss + LoRA grads vs the oracle composed layer by layer."""
    from llamax_b200.modelling import PrefixLM

    dynamic, P, B, S = True, 64, 2, 192
    model = build_tiny_llama(dynamic, num_layers=4)
    cfg = model.config
    torch.manual_seed(5)
    tokens = torch.randint(0, cfg.vocab_size, (B, S))
    labels = torch.randint(0, cfg.vocab_size, (B, S))
    labels[:, :P] = -100
    rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:S]

    # oracle: embed -> layers -> norm -> head -> CE, fp32
    lws = [oracle_layer_weights(l, torch.float32) for l in model.layers]
    x = model.tok_embeddings.weight.detach().float()[tokens]
    for lw in lws:
        x = R.transformer_layer_ref(x, rope, lw, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, P, dynamic)
    x = R.rmsnorm_ref(x, model.norm.weight.detach().float())
    logits = x @ model.output.weight.detach().float().T
    loss_ref = torch.nn.functional.cross_entropy(logits.view(-1, cfg.vocab_size), labels.view(-1))
    loss_ref.backward()

    model = model.cuda()
    model.build_cache()
    model.tok_embeddings.requires_grad_(False)
    model.output.requires_grad_(False)
    loss = model(tokens.cuda(), labels=labels.cuda(), block_mask=PrefixLM(P))
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 2e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    worst = 0.0
    for layer, lw in zip(model.layers, lws):
        att, ff = layer.attention, layer.feed_forward
        for name, mod in (("wq", att.wq), ("wo", att.wo), ("w1", ff.w1), ("w2", ff.w2)):
            worst = max(worst, rel_err(mod.lora_b.grad, lw.lora_b[name].grad))
    print("loss", loss.item(), loss_ref.item(), "worst lora_b grad err", worst)
    assert worst <= 5e-2


def test_fused_block_document_mask():
    """Packed-sequence document-causal mask (the mask the reference trainer ships) through the fused block."""
    from llamax_b200.modelling import DocumentCausal

    dynamic, B, S = True, 1, 448
    model = build_tiny_llama(dynamic, num_layers=1)
    layer, cfg = model.layers[0], model.config
    rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:S]
    doc_ids = torch.repeat_interleave(torch.arange(4), torch.tensor([130, 62, 200, 56]))[None]
    torch.manual_seed(9)
    x = torch.randn(B, S, cfg.embed_dim).bfloat16()
    dout = torch.randn(B, S, cfg.embed_dim).bfloat16()
    lw = oracle_layer_weights(layer, torch.float32)
    xr = x.float().requires_grad_(True)
    ref = R.transformer_layer_ref(xr, rope, lw, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, 0, dynamic, doc_ids=doc_ids)
    ref.backward(dout.float())
    layer = layer.cuda()
    xc = x.cuda().requires_grad_(True)
    out = layer(xc, rope.cuda(), block_mask=DocumentCausal(doc_ids.cuda()))
    out.backward(dout.cuda())
    assert rel_err(out, ref) <= 1.5e-2 and rel_err(xc.grad, xr.grad) <= 1.5e-2
    assert rel_err(layer.attention.wv.lora_b.grad, lw.lora_b["wv"].grad) <= 2e-2
    # documents do not leak: perturbing document 3 leaves the outputs of documents 0-2 unchanged
    x2 = x.clone()
    x2[:, 392:] += 1.0
    out2 = layer(x2.cuda(), rope.cuda(), block_mask=DocumentCausal(doc_ids.cuda()))
    assert torch.equal(out2[:, :392], out[:, :392])


def test_activation_checkpointing_matches_plain():
    """LlamaConfig.activation_checkpointing (llama.py:209-212) wraps the fused block in torch checkpoint: same loss and
    gradients as the plain run."""
    from llamax_b200.modelling import PrefixLM

    torch.manual_seed(13)
    tokens = torch.randint(0, 1024, (2, 160)).cuda()
    labels = torch.randint(0, 1024, (2, 160)).cuda()
    results = []
    for ckpt in (False, True):
        model = build_tiny_llama(True, num_layers=2)
        model.config = model.config._replace(activation_checkpointing=ckpt)
        model = model.cuda()
        model.build_cache()
        model.tok_embeddings.requires_grad_(False)
        model.output.requires_grad_(False)
        loss = model(tokens, labels=labels, block_mask=PrefixLM(40))
        loss.backward()
        results.append((loss.item(), model.layers[0].attention.wq.lora_b.grad.clone(), model.layers[1].ffn_norm.weight.grad.clone()))
    # fp32 reductions by atomics (loss sum, split-K LoRA gradients, dQ bulk-reduce) are order-dependent: compare closely
    assert abs(results[0][0] - results[1][0]) <= 1e-5 * abs(results[0][0])
    assert rel_err(results[1][1], results[0][1]) <= 2e-2 and rel_err(results[1][2], results[0][2]) <= 2e-2


def test_config0_tiny_audio_prefix_lm():
    """BASELINE.json configs[0] as stated: 4 layers, dim 512, GQA 8/2 (head_dim 64), audio Conv1D prefix, prefix-LM,
    seq 1024 (512 audio-prefix positions from 10.24 s of audio + 512 text), batch 2; weight-only INT8 + LoRA r=8.
    Loss and gradients (LoRA, conv stem) against the fp32 oracle composed from the same weights."""
    from llamax_b200.modelling import AudioConfig, LlamaAudio, LlamaConfig, apply_linear_adapter_
    from llamax_b200.subclasses import quantize_linear_

    torch.manual_seed(0)
    cfg = LlamaConfig(512, 4, 64, 8, 2, 1792, max_seq_len=1024, vocab_size=1024, rope_base=500000, is_llama3_1=True)
    model = LlamaAudio(cfg, AudioConfig(n_mels=80)).bfloat16()
    model.build_cache()
    dynamic = False
    quantize_linear_(model.layers, "int8", dynamic_int8_act=dynamic)
    apply_linear_adapter_(model.layers, "lora", rank=8)
    g = torch.Generator().manual_seed(1)
    for m in model.modules():
        if hasattr(m, "lora_b"):
            m.lora_b.data.copy_((torch.randn(m.lora_b.shape, generator=g) * 0.02).bfloat16())
    B, T = 2, 512
    audio = torch.randn(B, 163840, generator=g)          # 10.24 s -> 1025 frames -> 1024 -> 512 prefix positions
    tokens = torch.randint(0, 1024, (B, T), generator=g)
    labels = torch.randint(0, 1024, (B, T), generator=g)

    # oracle: the same front-end modules in fp32 on CPU, then the restated blocks, norm, head, CE
    import copy

    stem = copy.deepcopy(model.audio_embed).float()
    mel = model.melspec(audio)[..., :-1].clip(1e-12).log10()
    mel = (mel - mel.mean(2, keepdim=True)).bfloat16().float()
    prefix = stem(mel).transpose(1, 2)
    P = prefix.shape[1]
    assert P == 512
    x = torch.cat([prefix, model.tok_embeddings.weight.detach().float()[tokens]], 1)
    rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[: P + T]
    lws = [oracle_layer_weights(l, torch.float32) for l in model.layers]
    for lw in lws:
        x = R.transformer_layer_ref(x, rope, lw, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, P, dynamic)
    x = R.rmsnorm_ref(x[:, P:], model.norm.weight.detach().float())
    logits = x @ model.output.weight.detach().float().T
    loss_ref = torch.nn.functional.cross_entropy(logits.reshape(-1, 1024), labels.reshape(-1))
    loss_ref.backward()

    model = model.cuda()
    model.build_cache()
    model.tok_embeddings.requires_grad_(False)
    model.output.requires_grad_(False)
    loss = model(audio.cuda(), tokens.cuda(), labels=labels.cuda(), prefix_lm=True)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 2e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    errs = {
        "wq0.lora_b": rel_err(model.layers[0].attention.wq.lora_b.grad, lws[0].lora_b["wq"].grad),
        "w2_3.lora_a": rel_err(model.layers[3].feed_forward.w2.lora_a.grad, lws[3].lora_a["w2"].grad),
        "conv2.weight": rel_err(model.audio_embed[2].weight.grad, stem[2].weight.grad),
    }
    print(loss.item(), loss_ref.item(), errs)
    assert all(v <= 5e-2 for v in errs.values()), errs


def test_optin_int8_grad_input_is_close_but_not_parity():
    """SURVEY section 8 row f4 (opt-in, NON-parity): grad_input = q_rowwise(dY * s_w) @ W_int8 on the int8 tensor path.
    The forward is untouched (bit-identical output); gradients carry 8-bit row quantisation noise, so they are only
    required to stay within 5e-2 of the fp32 oracle — looser than the 1e-2 parity bar, which is why the mode is off by
    default and never used for the headline."""
    import llamax_b200.modelling.fused_block as FB

    base = _layer_case(True, 100)
    FB.set_int8_grad_input(True)
    try:
        rep = _layer_case(True, 100)
    finally:
        FB.set_int8_grad_input(False)
    print({k: (f"{rep[k][0]:.2e}", f"{base[k][0]:.2e}") for k in rep})
    assert rep["out"][0] == base["out"][0]
    for key, (err, _) in rep.items():
        assert err <= 5e-2, f"{key}: {err:.3e}"
    assert any(rep[k][0] != base[k][0] for k in ("dx", "a_wq", "an"))   # the mode really took the other path


@pytest.mark.parametrize("dynamic", [True, False])
def test_selective_recompute_matches_default(dynamic):
    """SURVEY section 8 row f4 (first half): with set_recompute("ffn") the block drops xn2 / w1x|w3x from its save-set and
    rebuilds them in backward: same output, gradients equal up to the run-to-run order of the fp32 L2 reductions, the
    rebuilt tensors bit-identical, and a smaller save-set."""
    import llamax_b200.modelling.fused_block as FB
    from llamax_b200.modelling import PrefixLM

    results, saved = [], []
    for policy in ("none", "ffn"):
        model = build_tiny_llama(dynamic, num_layers=1).cuda()
        layer, cfg = model.layers[0], model.config
        rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:256].cuda()
        torch.manual_seed(7)
        x = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16().requires_grad_(True)
        dout = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16()
        FB.set_recompute(policy)
        try:
            nbytes = 0

            def pack(t):
                nonlocal nbytes
                nbytes += t.numel() * t.element_size()
                return t

            with torch.autograd.graph.saved_tensors_hooks(pack, lambda t: t):
                out = layer(x, rope, block_mask=PrefixLM(64))
            out.backward(dout)
        finally:
            FB.set_recompute("none")
        grads = [x.grad] + [p.grad for p in layer.parameters() if p.requires_grad]
        results.append([out.detach()] + [g.detach().clone() for g in grads])
        saved.append(nbytes)
    assert len(results[0]) == len(results[1]) > 10
    assert torch.equal(results[0][0], results[1][0])                       # same forward
    for a, b in zip(results[0][1:], results[1][1:]):
        # the recomputed tensors are bit-identical (checked below); what differs run to run is only the summation
        # order of the fp32 reductions at L2 (dQ, LoRA dA/dB): a bf16 ulp here and there
        assert rel_err(a, b) <= 5e-3
    assert saved[1] < 0.6 * saved[0], saved
    # the rebuild itself: same kernels, same inputs -> the same bits, with or without the saved LoRA h
    model = build_tiny_llama(dynamic, num_layers=1).cuda()
    layer = model.layers[0]
    s1, s3 = FB.LinearSpec(layer.feed_forward.w1), FB.LinearSpec(layer.feed_forward.w3)
    x1 = torch.randn(512, model.config.embed_dim, device="cuda").bfloat16()
    w = layer.ffn_norm.weight.detach()
    xn_a, rs_a, ab_a, h = FB._ffn_up(x1, w, s1, s3, s1.dynamic, None)
    xn_b, rs_b, ab_b, _ = FB._ffn_up(x1, w, s1, s3, s1.dynamic, h)
    assert torch.equal(xn_a, xn_b) and torch.equal(rs_a, rs_b) and torch.equal(ab_a, ab_b)


def test_fused_block_lora_rank16():
    """rank 16 on all seven linears: the q|k|v group carries 48 LoRA columns (epilogue rank 16 per GEMM, dh columns
    K-concatenated into the grad_input GEMM, dA through the 32-column wgrad kernel in two chunks)."""
    report = _layer_case(True, 64, rank=16)
    check_parity_table(report, "fused block dim 512, rank 16, P=64", KNOWN_BF16_LIMITED)


@pytest.mark.parametrize("adapters", ["attention", "none"])
def test_fused_block_partial_or_no_adapters(adapters):
    """LoRA only on the attention projections, or no adapters at all (frozen INT8 block, only the norms train)."""
    report = _layer_case(True, 0, adapters=adapters)
    assert ("a_w1" not in report) and (("a_wq" in report) == (adapters == "attention"))
    check_parity_table(report, "fused block dim 512 (True, 0, adapters=adapters)", KNOWN_BF16_LIMITED)


def test_fused_block_ragged_token_count():
    """B*S = 301 tokens: not a multiple of any tile (128-row GEMM / attention tiles, 8-element TMA pitches of the
    transposed LoRA operands)."""
    report = _layer_case(True, 77, B=1, S=301)
    check_parity_table(report, "fused block dim 512 (True, 77, B=1, S=301)", KNOWN_BF16_LIMITED)


def test_trainable_lm_head_loss_and_weight_gradient():
    """LM head + cross-entropy (llama.py:216-218) with a TRAINABLE head: loss, dx and dW (the dW chunks run on the
    weight-gradient GEMM form, dlogits^T x with both operands as stored) against the fp32 evaluation. Ragged row count
    (two chunks, the second partial), -100 labels."""
    from llamax_b200.modelling.llama import _ChunkedLMLoss, chunked_lm_loss

    torch.manual_seed(5)
    M, D, V = 700, 256, 1032
    x = (torch.randn(M, D) * 0.5).bfloat16()
    w = (torch.randn(V, D) * 0.05).bfloat16()
    labels = torch.randint(0, V, (M,))
    labels[::7] = -100
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    loss_ref = torch.nn.functional.cross_entropy(xr @ wr.t(), labels)
    loss_ref.backward()
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    old = _ChunkedLMLoss.CHUNK
    _ChunkedLMLoss.CHUNK = 512
    try:
        loss = chunked_lm_loss(xc, wc, labels.cuda(), wc.detach().t().contiguous())
        loss.backward()
    finally:
        _ChunkedLMLoss.CHUNK = old
    assert abs(loss.item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item())
    assert rel_err(xc.grad, xr.grad) <= TOL
    assert rel_err(wc.grad, wr.grad) <= TOL


def test_training_trajectory_follows_oracle():
    """Six optimisation steps on a fixed batch (2-layer model, prefix-LM, dynamic INT8 + LoRA): the product path's loss
    trajectory must follow the oracle's (same update rule on both sides: a sign-free normalised gradient step on the
    LoRA matrices and norm weights), step by step within 5e-3 (measured: 3e-4), and the loss must go down. Catches anything a single
    forward/backward cannot: gradients written to the wrong parameter, stale cached operands after an update (the
    resident (scale*W)^T | A^T buffers of the block backward), accumulation across steps."""
    from llamax_b200.modelling import PrefixLM

    dynamic, P, B, S, steps, lr = True, 48, 2, 160, 6, 0.05
    model = build_tiny_llama(dynamic, num_layers=2)
    cfg = model.config
    torch.manual_seed(9)
    tokens = torch.randint(0, cfg.vocab_size, (B, S))
    labels = torch.randint(0, cfg.vocab_size, (B, S))
    labels[:, :P] = -100
    rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:S]
    lws = [oracle_layer_weights(l, torch.float32) for l in model.layers]
    emb = model.tok_embeddings.weight.detach().float()
    w_norm, w_out = model.norm.weight.detach().float(), model.output.weight.detach().float()

    def update(params):
        with torch.no_grad():
            for p_ in params:
                if p_.grad is not None:
                    p_ -= (lr * p_.grad.float() / p_.grad.float().abs().max().clamp_min(1e-12)).to(p_.dtype)
                    p_.grad = None

    ref_params = [t for lw in lws for t in (*lw.lora_a.values(), *lw.lora_b.values(), lw.attention_norm, lw.ffn_norm)]
    ref_losses = []
    for _ in range(steps):
        x = emb[tokens]
        for lw in lws:
            x = R.transformer_layer_ref(x, rope, lw, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, P, dynamic)
        logits = R.rmsnorm_ref(x, w_norm) @ w_out.T
        loss = torch.nn.functional.cross_entropy(logits.view(-1, cfg.vocab_size), labels.view(-1))
        loss.backward()
        ref_losses.append(loss.item())
        update(ref_params)

    model = model.cuda()
    model.build_cache()
    for m in (model.tok_embeddings, model.output, model.norm):
        m.requires_grad_(False)
    params = [p_ for p_ in model.parameters() if p_.requires_grad]
    assert len(params) == len(ref_params)
    losses = []
    tk, lb = tokens.cuda(), labels.cuda()
    for _ in range(steps):
        loss = model(tk, labels=lb, block_mask=PrefixLM(P))
        loss.backward()
        losses.append(loss.item())
        update(params)
    print("loss trajectory ours", [f"{v:.4f}" for v in losses], "oracle", [f"{v:.4f}" for v in ref_losses])
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 5e-3 * abs(b), (losses, ref_losses)
    assert losses[-1] < losses[0] - 0.05 and ref_losses[-1] < ref_losses[0] - 0.05


def test_weight_cache_modes_give_identical_gradients():
    """The grad_input operand (scale * W)^T is either resident in bf16 ("1": 2 extra bytes per parameter, 14 GB at 8B) or
    rebuilt per use from the int8 codes into one shared scratch ("0": no extra weight memory, +2 % step time — profiles/
    r2_bench_1gpu_weight_cache_off.json): the operand holds the same bf16 numbers either way, so the block output is
    bit-identical and the gradients agree to the fp32-reduction-order noise."""
    from llamax_b200.modelling import PrefixLM
    from llamax_b200.modelling import fused_block as FB

    results = []
    for mode in ("1", "0"):
        FB.set_weight_cache(mode)
        try:
            model = build_tiny_llama(True, num_layers=1).cuda()
            layer, cfg = model.layers[0], model.config
            rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:256].cuda()
            torch.manual_seed(7)
            x = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16().requires_grad_(True)
            dout = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16()
            for _ in range(2):      # second pass: the resident operands are hits
                x.grad = None
                for p_ in layer.parameters():
                    p_.grad = None
                out = layer(x, rope, block_mask=PrefixLM(64))
                out.backward(dout)
            results.append([out.detach(), x.grad.detach()] + [p_.grad.detach() for p_ in layer.parameters() if p_.requires_grad])
        finally:
            FB.set_weight_cache("auto")
    for other in results[1:]:
        assert torch.equal(results[0][0], other[0])
        for a, b in zip(results[0][1:], other[1:]):
            assert rel_err(a, b) <= 5e-3


@pytest.mark.parametrize("dynamic", [True, False])
def test_mixed_input_gemm_mode_gives_identical_block(dynamic):
    """LLAMAX_MIXED_GEMM=1: the weight-only forward and every grad_input GEMM read the frozen INT8 weights
    directly (expanded to bf16 inside the GEMM) instead of a de-quantised bf16 operand. The expansion is exact and the
    scale product is rounded the same way, so the block output and the input gradient are BIT-identical to the default
    path in both INT8 modes (the parameter gradients only see fp32 reduction-order noise)."""
    from llamax_b200.modelling import PrefixLM
    from llamax_b200.modelling import fused_block as FB

    results = []
    for mixed in (False, True):
        FB.set_mixed_gemm(mixed)
        try:
            model = build_tiny_llama(dynamic, num_layers=1).cuda()
            layer, cfg = model.layers[0], model.config
            rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:256].cuda()
            torch.manual_seed(7)
            x = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16().requires_grad_(True)
            dout = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16()
            for _ in range(2):
                x.grad = None
                for p_ in layer.parameters():
                    p_.grad = None
                out = layer(x, rope, block_mask=PrefixLM(64))
                out.backward(dout)
            if mixed:   # no bf16 grad_input operand was built
                assert not any(k in layer.__dict__.get("_llamax_bwd_operands", {}) for k in ("w2", "w13", "wo", "wqkv"))
            results.append([out.detach(), x.grad.detach()] + [p_.grad.detach() for p_ in layer.parameters() if p_.requires_grad])
        finally:
            FB.set_mixed_gemm(False)
    assert torch.equal(results[0][0], results[1][0])
    assert rel_err(results[1][1], results[0][1]) <= 5e-3   # dQ is reduced with fp32 atomics: a bf16 ulp here and there
    for a, b in zip(results[0][2:], results[1][2:]):
        assert rel_err(b, a) <= 5e-3


def test_lm_head_row_compaction_changes_nothing():
    """Rows whose label is -100 are dropped before the final norm / LM head / cross-entropy (llamax_b200.modelling.llama
    _LM_COMPACT): loss and every gradient equal the all-rows computation (the dropped rows' contributions are exact
    zeros there), also with ignored positions scattered through the batch and with a whole sequence ignored."""
    from llamax_b200.modelling import PrefixLM
    from llamax_b200.modelling import llama as L

    B, S = 3, 256
    model = build_tiny_llama(True, num_layers=2).cuda()
    model.build_cache()
    model.tok_embeddings.requires_grad_(False)
    model.output.requires_grad_(False)
    cfg = model.config
    torch.manual_seed(11)
    tokens = torch.randint(0, cfg.vocab_size, (B, S), device="cuda")
    labels = torch.randint(0, cfg.vocab_size, (B, S), device="cuda")
    labels[:, :70] = -100
    labels[torch.rand(B, S, device="cuda") < 0.2] = -100
    labels[2] = -100
    results = []
    for on in (False, True):
        L.set_lm_compact(on)
        try:
            for p_ in model.parameters():
                p_.grad = None
            loss = model(tokens, labels=labels, block_mask=PrefixLM(64))
            loss.backward()
            results.append([loss.detach().float()] + [p_.grad.detach().float().clone() for p_ in model.parameters() if p_.grad is not None])
        finally:
            L.set_lm_compact(True)
    assert len(results[0]) == len(results[1]) > 10
    assert abs(results[0][0].item() - results[1][0].item()) <= 1e-5 * abs(results[0][0].item())
    for a, b in zip(results[0][1:], results[1][1:]):
        assert rel_err(b, a) <= 5e-3   # fp32-atomic reduction order of dQ / LoRA gradients only
    # every row labelled / no row labelled: the all-rows path
    full = torch.randint(0, cfg.vocab_size, (B, S), device="cuda")
    assert torch.isfinite(model(tokens, labels=full, block_mask=PrefixLM(64)))
    none = torch.full((B, S), -100, device="cuda")
    assert model(tokens, labels=none, block_mask=PrefixLM(64)).item() == 0.0


def test_delta_from_gemm_epilogue_mode_matches_default():
    """LLAMAX_FUSE_DELTA=1: delta of the attention backward from the wo grad_input GEMM's epilogue (head_dim 128)."""
    from llamax_b200.modelling import PrefixLM
    from llamax_b200.modelling import fused_block as FB

    results = []
    for on in (False, True):
        FB.set_fuse_delta(on)
        try:
            model = build_tiny_llama(True, num_layers=1).cuda()
            layer, cfg = model.layers[0], model.config
            rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:256].cuda()
            torch.manual_seed(7)
            x = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16().requires_grad_(True)
            dout = torch.randn(2, 256, cfg.embed_dim, device="cuda").bfloat16()
            out = layer(x, rope, block_mask=PrefixLM(64))
            out.backward(dout)
            results.append([x.grad.detach()] + [p_.grad.detach() for p_ in layer.parameters() if p_.requires_grad])
        finally:
            FB.set_fuse_delta(False)
    for a, b in zip(results[0], results[1]):
        assert rel_err(b, a) <= 5e-3


@pytest.mark.parametrize("dynamic", [True, False])
def test_grouped_lora_backward_equals_one_launch_per_adapter(dynamic, monkeypatch):
    """wq|wk|wv and w1|w3 take ONE lora_bwd_pair launch each over a block-diagonal scale*B^T (fused_block._LORA_GROUP_PAIR):
    dh (hence the input gradient's LoRA term) must be the bits of the per-adapter launches — the off-diagonal zeros add
    exact zeros — and every dB the diagonal block of the merged result (equal up to the order of the fp32 L2 reductions)."""
    import llamax_b200.modelling.fused_block as FB
    from llamax_b200.modelling import PrefixLM

    results = []
    for grouped in (True, False):
        monkeypatch.setattr(FB, "_LORA_GROUP_PAIR", grouped)
        model = build_tiny_llama(dynamic, num_layers=1).cuda()      # helpers give every lora_b non-zero values
        layer, cfg = model.layers[0], model.config
        rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:512].cuda()
        torch.manual_seed(5)
        x = torch.randn(2, 512, cfg.embed_dim, device="cuda").bfloat16().requires_grad_(True)
        dout = torch.randn(2, 512, cfg.embed_dim, device="cuda").bfloat16()
        out = layer(x, rope, block_mask=PrefixLM(300))
        out.backward(dout)
        names = [n for n, p in layer.named_parameters() if p.requires_grad]
        results.append((out.detach(), x.grad.detach().clone(),
                        {n: p.grad.detach().clone() for n, p in layer.named_parameters() if p.requires_grad}))
        assert any(n.endswith("wk.lora_b") for n in names) and any(n.endswith("w3.lora_a") for n in names)
    (o1, dx1, g1), (o0, dx0, g0) = results
    assert torch.equal(o1, o0)
    assert rel_err(dx1, dx0) <= 5e-3          # dQ is reduced with fp32 atomics: a bf16 ulp here and there, run to run
    assert g1.keys() == g0.keys()
    for n in g1:
        assert g1[n].shape == g0[n].shape and rel_err(g1[n], g0[n]) <= 5e-3, n


def test_rope_in_qkv_epilogue_changes_nothing(monkeypatch):
    """fused_block._ROPE_EPILOGUE (opt-in): the block's output and gradients with RoPE applied by the q | k | v GEMM's
    epilogue equal those with the separate in-place pass (head_dim 128: the 8B-shape attention at reduced width)."""
    import llamax_b200.modelling.fused_block as FB
    from llamax_b200.modelling import PrefixLM
    from llamax_b200.modelling.llama import LlamaConfig

    cfg = LlamaConfig(1024, 1, 128, 8, 2, 2048, max_seq_len=512, vocab_size=512, rope_base=500000, is_llama3_1=True)
    results = []
    for fused in (True, False):
        monkeypatch.setattr(FB, "_ROPE_EPILOGUE", fused)
        model = build_tiny_llama(True, num_layers=1, config=cfg).cuda()
        layer = model.layers[0]
        rope = R.build_rope(cfg.head_dim, cfg.max_seq_len, cfg.rope_base, cfg.is_llama3_1)[:384].contiguous().cuda()
        torch.manual_seed(3)
        x = torch.randn(2, 384, cfg.embed_dim, device="cuda").bfloat16().requires_grad_(True)
        dout = torch.randn(2, 384, cfg.embed_dim, device="cuda").bfloat16()
        out = layer(x, rope, block_mask=PrefixLM(100))
        out.backward(dout)
        results.append((out.detach(), x.grad.detach().clone(),
                        [p.grad.detach().clone() for p in layer.parameters() if p.requires_grad]))
    (o1, dx1, g1), (o0, dx0, g0) = results
    assert torch.equal(o1, o0)
    assert rel_err(dx1, dx0) <= 5e-3 and len(g1) == len(g0) > 10
    for a, b in zip(g1, g0):
        assert rel_err(a, b) <= 5e-3
