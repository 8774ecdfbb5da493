"""llamax_b200 — B200-native (sm_100a) fine-tuning hot path of gau-nernst/llama-x.

Drop-in packages mirroring the reference layout: `llamax_b200.subclasses` (INT8 tensor subclass, int8_mm_dequant
op) and `llamax_b200.modelling` (Llama / LlamaAudio / LoRALinear). The compute lives in
`csrc/libllamax_b200.so` (C ABI in include/llamax_b200.h); there is no CPU or eager fallback.
"""

__version__ = "0.1.0"
