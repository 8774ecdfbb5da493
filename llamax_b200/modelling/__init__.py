from .audio import AudioConfig, LlamaAudio
from .llama import (DocumentCausal, Llama, LlamaConfig, PrefixLM, document_block_mask, prefix_lm_attention,
                    prefix_lm_block_mask)
from .lora import LoRALinear, apply_linear_adapter_
