"""LoRA adapters on (quantised) linear layers — drop-in for modelling/lora.py of the reference.

`apply_linear_adapter_(model, "lora", rank=8, alpha=8.0)` class-swaps every nn.Linear to LoRALinear exactly like the
reference (lora.py:8-16), parameter names `lora_a [rank, in]`, `lora_b [out, rank]`, init kaiming-normal(a=sqrt 5) /
zeros, scale = alpha / rank (lora.py:20-35).  `LoRALinear.forward` computes
    F.linear(x, W, b) + (x @ A^T) @ B^T * scale                                             (lora.py:40-44)
With an Int8LinearWeight base on CUDA this is ONE fused op: the LoRA up-projection and the scale ride in the
epilogue of the base GEMM, and the backward shares the de-quantised operand with grad_input.
DoRA is not part of this path (it does not compose with the INT8 subclass in the reference either).
"""

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .. import ops
from ..subclasses.int8 import (Int8LinearWeight, _require_cuda_bf16, int8_linear_forward, int8_linear_grad_input,
                               quantize_int8_rowwise)


def apply_linear_adapter_(model: nn.Module, adapter: str | None, **kwargs):
    if adapter is None:
        return
    if adapter != "lora":
        raise NotImplementedError(f"adapter {adapter!r}: only 'lora' is implemented on the B200 path")
    for m in model.modules():
        if isinstance(m, nn.Linear):
            m.__class__ = LoRALinear
            m.init_adapter(**kwargs)


class _LoRAInt8Linear(torch.autograd.Function):
    """y = int8_linear(x, W) + scale * (x A^T) B^T with W frozen."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Int8LinearWeight, lora_a: Tensor, lora_b: Tensor, scale: float):
        _require_cuda_bf16(x, "LoRALinear.forward")
        x2 = x.reshape(-1, weight.shape[1])
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        h = ops.bf16_gemm(x2, lora_a.detach())  # [M, r]
        y2 = int8_linear_forward(x2, weight.int_data, weight.scale, weight.dynamic_int8_act,
                                 lora_h=h, lora_b=lora_b.detach(), lora_scale=scale)
        ctx.save_for_backward(x2, h, weight.int_data, weight.scale, lora_a, lora_b)
        ctx.scale = scale
        return y2.view(*x.shape[:-1], -1)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        x2, h, w8, w_scale, lora_a, lora_b = ctx.saved_tensors
        scale = ctx.scale
        g2 = grad_out.reshape(-1, w8.shape[0])
        if g2.stride(1) != 1:
            g2 = g2.contiguous()
        dh = ops.bf16_gemm(g2, (lora_b.detach().t() * scale).contiguous())  # [M, r] = scale * dy B
        dx = dA = dB = None
        if ctx.needs_input_grad[0]:
            dx = int8_linear_grad_input(g2, w8, w_scale, lora_h=dh, lora_b=lora_a.detach().t().contiguous(),
                                        lora_scale=1.0).view(*grad_out.shape[:-1], -1)
        if ctx.needs_input_grad[2]:
            dA = ops.lora_wgrad(x2, dh, 1.0).t().to(lora_a.dtype).contiguous()
        if ctx.needs_input_grad[3]:
            dB = ops.lora_wgrad(g2, h, scale).to(lora_b.dtype)
        return dx, None, dA, dB, None


class LoRALinear(nn.Linear):
    def init_adapter(self, rank: int = 8, alpha: float = 8.0) -> None:
        self.weight.requires_grad_(False)
        if self.bias is not None:
            self.bias.requires_grad_(False)
        self.rank, self.alpha = rank, alpha
        self.scale = alpha / rank if rank > 0 else 0.0
        if rank > 0:
            kw = dict(dtype=self.weight.dtype, device=self.weight.device)
            self.lora_a = nn.Parameter(torch.empty(rank, self.in_features, **kw))
            self.lora_b = nn.Parameter(torch.zeros(self.out_features, rank, **kw))
            nn.init.kaiming_normal_(self.lora_a, a=5**0.5)

    def extra_repr(self):
        return f"{super().extra_repr()}, rank={self.rank}, alpha={self.alpha}"

    def forward(self, x: Tensor):
        fused = (self.rank > 0 and self.rank % 8 == 0 and self.rank <= 16 and self.bias is None
                 and isinstance(self.weight, Int8LinearWeight) and x.is_cuda)
        if fused:
            return _LoRAInt8Linear.apply(x, self.weight, self.lora_a, self.lora_b, self.scale)
        out = F.linear(x, self.weight, self.bias)
        if self.rank > 0:
            out = out + x @ self.lora_a.T @ self.lora_b.T * self.scale
        return out
