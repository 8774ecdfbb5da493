"""INT8 tensor subclass for frozen linear weights: the `subclasses.int8` seam of llama-x, B200 backend.

Interface contract (what callers of the reference rely on, reference file:line in brackets):
  * quantize_int8_rowwise(x) -> (int8 codes, scale in x.dtype)                          [subclasses/int8.py:10-16]
  * Int8LinearWeight(int_data, scale, dynamic_int8_act=False): wrapper subclass whose outer dtype is the scale
    dtype; attributes .int_data [N,K] int8, .scale [N], .dynamic_int8_act; from_float(); dequantize();
    flatten/unflatten protocol; intercepts F.linear; supports detach / clone / .to() / copy_ and nothing else
                                                                                       [subclasses/int8.py:19-102]
  * _Int8Linear.apply(input, weight, bias) with bias positional; backward returns (grad_input, None, grad_bias)
                                                                                       [subclasses/int8.py:106-130]

Backend: F.linear runs on this package's sm_100a kernels — fused row-quantisation, tcgen05 INT8 GEMM with the
row/column-scale dequant in the epilogue (dynamic mode), or a bf16 tcgen05 GEMM on a de-quantised weight operand
(weight-only forward, and grad_input in both modes). CUDA + bfloat16 only: any other device or dtype raises,
there is no eager fallback.

Quantising checkpoint weights on the host before `.cuda()` (what the train scripts do once at start-up) works
through ordinary tensor ops: that is module surgery, not part of the training step.
"""

import torch
import torch.nn.functional as F
from torch import Tensor

from .. import ops
from .int8_mm import int8_mm_dequant

_aten = torch.ops.aten


def quantize_int8_rowwise(x: Tensor):
    if x.is_cuda and x.dtype is torch.bfloat16 and x.dim() == 2:
        return ops.rowquant_int8(x)  # one fused pass: amax, IEEE divide, round-half-even
    # host-side setup path (weights at model-preparation time); same arithmetic
    as_f32 = x.to(torch.float32)
    row_scale = as_f32.abs().amax(dim=1) / 127
    codes = torch.round(as_f32 / row_scale.clamp(min=1e-12)[:, None]).to(torch.int8)
    return codes, row_scale.to(x.dtype)


def _rewrap_unary(cls, func, args, kwargs):
    w = args[0]
    rest = args[1:]
    return cls(func(w.int_data, *rest, **kwargs), func(w.scale, *rest, **kwargs), w.dynamic_int8_act)


def _to_copy(cls, func, args, kwargs):
    w = args[0]
    device, dtype = kwargs.get("device"), kwargs.get("dtype")
    return cls(w.int_data.to(device=device), w.scale.to(device=device, dtype=dtype), w.dynamic_int8_act)


def _copy_(cls, func, args, kwargs):
    dst, src = args[0], args[1]
    dst_q, src_q = isinstance(dst, cls), isinstance(src, cls)
    if dst_q and src_q:
        dst.int_data.copy_(src.int_data)
        dst.scale.copy_(src.scale)
    elif dst_q:  # float -> quantised: re-quantise
        codes, row_scale = quantize_int8_rowwise(src)
        dst.int_data.copy_(codes)
        dst.scale.copy_(row_scale)
    else:  # quantised -> float
        dst.copy_(src.dequantize())
    return dst


_DISPATCH = {
    _aten.detach.default: _rewrap_unary,
    _aten.clone.default: _rewrap_unary,
    _aten._to_copy.default: _to_copy,
    _aten.copy_.default: _copy_,
}


class Int8LinearWeight(Tensor):
    @staticmethod
    @torch._dynamo.disable
    def __new__(cls, int_data: Tensor, scale: Tensor, dynamic_int8_act: bool = False):
        return Tensor._make_wrapper_subclass(cls, int_data.shape, dtype=scale.dtype, device=int_data.device)

    @torch._dynamo.disable
    def __init__(self, int_data: Tensor, scale: Tensor, dynamic_int8_act: bool = False):
        assert int_data.dtype is torch.int8 and int_data.ndim == 2, "int_data must be a 2-D int8 tensor"
        assert scale.ndim == 1, "scale must be 1-D (one entry per output row)"
        self.int_data = int_data
        self.scale = scale
        self.dynamic_int8_act = dynamic_int8_act

    # -- traceable-subclass protocol ---------------------------------------------------------------
    def __tensor_flatten__(self):
        return ["int_data", "scale"], [self.dynamic_int8_act]

    @classmethod
    def __tensor_unflatten__(cls, tensor_data_dict, tensor_attributes, outer_size=None, outer_stride=None):
        return cls(tensor_data_dict["int_data"], tensor_data_dict["scale"], *tensor_attributes)

    # -- construction / inspection -----------------------------------------------------------------
    @classmethod
    def from_float(cls, tensor: Tensor, dynamic_int8_act: bool = False):
        codes, row_scale = quantize_int8_rowwise(tensor)
        return cls(codes, row_scale, dynamic_int8_act)

    def dequantize(self):
        return self.int_data * self.scale.view(-1, 1)

    def __repr__(self):
        return (
            f"{type(self).__name__}(shape={tuple(self.shape)}, dynamic_int8_act={self.dynamic_int8_act}, "
            f"dtype={self.dtype}, device={self.device}, requires_grad={self.requires_grad})"
        )

    # -- interception ------------------------------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        if func is F.linear:
            return _Int8Linear.apply(*args, **(kwargs or {}))
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **(kwargs or {}))

    @classmethod
    def __torch_dispatch__(cls, func, types, args, kwargs):
        handler = _DISPATCH.get(func)
        if handler is None:
            raise NotImplementedError(f"{cls.__name__} dispatch: attempting to run {func}, this is not supported")
        return handler(cls, func, args, kwargs or {})


def _require_cuda_bf16(t: Tensor, what: str):
    if not t.is_cuda or t.dtype is not torch.bfloat16:
        raise NotImplementedError(
            f"llamax_b200 {what}: only CUDA bfloat16 is implemented (got {t.device}, {t.dtype}); "
            "this package has no CPU or eager-PyTorch fallback for the hot path"
        )


def int8_linear_forward(x2: Tensor, int_data: Tensor, scale: Tensor, dynamic: bool, **epilogue) -> Tensor:
    """x2 [M,K] bf16 -> [M,N] bf16 in either INT8 mode; `epilogue`: lora_h / lora_b / lora_scale / resid / out."""
    if dynamic:
        x8, xs = ops.rowquant_int8(x2)
        return ops.int8_gemm_dequant(x8, int_data, xs, scale, **epilogue)
    # weight-only: bf16(bf16(x @ W8^T) * s), both roundings of the reference kept (int8.py:118)
    w_bf16 = ops.dequant_weight(int_data, None, transpose=False, apply_scale=False)
    return ops.bf16_gemm(x2, w_bf16, col_scale=scale, round_before_scale=True, **epilogue)


def int8_linear_grad_input(dy2: Tensor, int_data: Tensor, scale: Tensor, **epilogue) -> Tensor:
    """dy2 [M,N] -> dx [M,K] = dy @ (s * W8) (int8.py:127). The row scale sits on the contraction index, so it is
    folded into the de-quantised operand instead of a separate pass over dy."""
    w_t = ops.dequant_weight(int_data, scale, transpose=True, apply_scale=True)  # [K, N]
    return ops.bf16_gemm(dy2, w_t, **epilogue)


class _Int8Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input: Tensor, weight: Int8LinearWeight, bias: Tensor | None = None):
        _require_cuda_bf16(input, "F.linear(Int8LinearWeight)")
        ctx.save_for_backward(weight.int_data, weight.scale)  # frozen base: the input is not needed
        ctx.has_bias = bias is not None
        x2 = input.reshape(-1, weight.shape[1])
        if weight.dynamic_int8_act:
            # same seams as the reference: quantize_int8_rowwise -> int8_mm_dequant(A, W8.T, sA, sW)
            x8, xs = quantize_int8_rowwise(x2)
            y2 = int8_mm_dequant(x8, weight.int_data.T, xs, weight.scale)
        else:
            y2 = int8_linear_forward(x2, weight.int_data, weight.scale, False)
        y = y2.view(*input.shape[:-1], -1)
        return y if bias is None else y + bias

    @staticmethod
    def backward(ctx, grad_output: Tensor):
        w8, w_scale = ctx.saved_tensors
        g2 = grad_output.reshape(-1, w8.shape[0])
        if g2.stride(1) != 1:
            g2 = g2.contiguous()
        grad_input = grad_bias = None
        if ctx.needs_input_grad[0]:
            grad_input = int8_linear_grad_input(g2, w8, w_scale).view(*grad_output.shape[:-1], -1)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            grad_bias = g2.sum(0)
        return grad_input, None, grad_bias
