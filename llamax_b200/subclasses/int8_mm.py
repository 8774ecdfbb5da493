"""`int8_mm_dequant(A, B, A_scale, B_scale) -> Tensor` — operator seam of the reference
(subclasses/int8_mm.py:121-149), backed by the tcgen05 INT8 GEMM instead of the Triton kernel (:50-118).

Same schema, same Python-level asserts, same output convention (fresh contiguous [M, N] tensor in the dtype of
A_scale). The op lives in this package's own library namespace (`llamax_b200::int8_mm_dequant`) so that it cannot
clash with a real torchao install; it has a Meta implementation (shape inference for tracing) and a CUDA
implementation. Like the reference, there is no CPU implementation.
"""

import torch
from torch import Tensor

from .. import ops

lib = torch.library.Library("llamax_b200", "FRAGMENT")
lib.define("int8_mm_dequant(Tensor A, Tensor B, Tensor A_scale, Tensor B_scale) -> Tensor")


def int8_mm_dequant(A: Tensor, B: Tensor, A_scale_rowwise: Tensor, B_scale_colwise: Tensor) -> Tensor:
    assert A.dtype is torch.int8 and B.dtype is torch.int8
    assert A_scale_rowwise.dtype is B_scale_colwise.dtype
    assert A.shape[1] == B.shape[0]
    assert A_scale_rowwise.squeeze().shape == (A.shape[0],)
    assert B_scale_colwise.squeeze().shape == (B.shape[1],)
    assert A_scale_rowwise.is_contiguous()
    assert B_scale_colwise.is_contiguous()
    return torch.ops.llamax_b200.int8_mm_dequant(A, B, A_scale_rowwise, B_scale_colwise)


@torch.library.impl(lib, "int8_mm_dequant", "Meta")
def _(A: Tensor, B: Tensor, A_scale_rowwise: Tensor, B_scale_colwise: Tensor):
    return torch.empty((A.shape[0], B.shape[1]), device=A.device, dtype=A_scale_rowwise.dtype)


@torch.library.impl(lib, "int8_mm_dequant", "CUDA")
def int8_mm_dequant_cuda(A: Tensor, B: Tensor, A_scale_rowwise: Tensor, B_scale_colwise: Tensor):
    if A_scale_rowwise.dtype is not torch.bfloat16:
        raise NotImplementedError("llamax_b200::int8_mm_dequant: only bf16 scales/outputs are implemented")
    # The kernel wants both operands contraction-contiguous: A [M,K] and B^T [N,K]. The reference passes
    # B = weight.int_data.T (strides (1, K)), for which B.T is already that layout (no copy).
    if A.stride(1) != 1:
        A = A.contiguous()
    Bt = B.T
    if Bt.stride(1) != 1:
        Bt = Bt.contiguous()
    return ops.int8_gemm_dequant(A, Bt, A_scale_rowwise, B_scale_colwise)
