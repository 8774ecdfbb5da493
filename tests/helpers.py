"""Shared builders for the parity tests: a tiny model on the product path and the matching oracle weights."""
import torch

from oracle import ref_ops as R


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| / max |b| (the tolerance metric stated in BASELINE.json's north_star)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def tiny_config(num_layers=2, max_seq_len=512):
    from llamax_b200.modelling import LlamaConfig

    # dim 512, GQA 4/1 with head_dim 128 (the 8B head shape), ffn 1792, Llama-3.1 RoPE
    return LlamaConfig(512, num_layers, 128, 4, 1, 1792, max_seq_len=max_seq_len, vocab_size=1024,
                       rope_base=500000, is_llama3_1=True)


def build_tiny_llama(dynamic: bool, num_layers=2, rank=8, seed=0, audio=False, max_seq_len=512, adapters="all"):
    """CPU-initialised (deterministic), quantised + LoRA'd model; caller moves it to CUDA.
    adapters: "all" | "attention" (LoRA on wq/wk/wv/wo only) | "none"."""
    from llamax_b200.modelling import AudioConfig, Llama, LlamaAudio, apply_linear_adapter_
    from llamax_b200.subclasses import quantize_linear_

    torch.manual_seed(seed)
    cfg = tiny_config(num_layers, max_seq_len)
    model = LlamaAudio(cfg, AudioConfig(n_mels=80)) if audio else Llama(cfg)
    model = model.bfloat16()
    quantize_linear_(model.layers, "int8", dynamic_int8_act=dynamic)
    if adapters == "all":
        apply_linear_adapter_(model.layers, "lora", rank=rank)
    elif adapters == "attention":
        for layer in model.layers:
            apply_linear_adapter_(layer.attention, "lora", rank=rank)
    g = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():
        if hasattr(m, "lora_b"):  # zeros-init hides dA / dx_lora errors
            m.lora_b.data.copy_((torch.randn(m.lora_b.shape, generator=g) * 0.02).bfloat16())
    for layer in model.layers:
        for n in (layer.attention_norm, layer.ffn_norm):
            n.weight.data.copy_((1 + 0.1 * torch.randn(n.weight.shape, generator=g)).bfloat16())
    return model


def oracle_layer_weights(layer, dtype=torch.bfloat16) -> R.LayerWeights:
    """Copy one TransformerLayer's tensors into the oracle's container (CPU; float leaves require grad)."""
    lw = R.LayerWeights()
    att, ff = layer.attention, layer.feed_forward
    for name, mod in (("wq", att.wq), ("wk", att.wk), ("wv", att.wv), ("wo", att.wo), ("w1", ff.w1), ("w3", ff.w3),
                      ("w2", ff.w2)):
        lw.w8[name] = mod.weight.int_data.detach().cpu()
        lw.ws[name] = mod.weight.scale.detach().cpu().to(dtype)
        if getattr(mod, "rank", 0) > 0:
            lw.lora_a[name] = mod.lora_a.detach().cpu().to(dtype).requires_grad_(True)
            lw.lora_b[name] = mod.lora_b.detach().cpu().to(dtype).requires_grad_(True)
            lw.lora_scale = mod.scale
    lw.attention_norm = layer.attention_norm.weight.detach().cpu().to(dtype).requires_grad_(True)
    lw.ffn_norm = layer.ffn_norm.weight.detach().cpu().to(dtype).requires_grad_(True)
    return lw
