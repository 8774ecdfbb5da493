"""Drop-in for the reference's `subclasses` package (subclasses/__init__.py:1-13)."""

from torch import nn

from .int8 import Int8LinearWeight, quantize_int8_rowwise
from .int8_mm import int8_mm_dequant


def quantize_linear_(model: nn.Module, quantize: str | None, **kwargs):
    """Replace the weight of every nn.Linear in `model` by a frozen quantised parameter (in place)."""
    if quantize is None:
        return
    fn = dict(int8=Int8LinearWeight.from_float)[quantize]
    for m in model.modules():
        if isinstance(m, nn.Linear):
            m.weight = nn.Parameter(fn(m.weight.detach(), **kwargs), requires_grad=False)
