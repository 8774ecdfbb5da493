"""Loader for the UNMODIFIED reference installed under the git-ignored `baseline/_ref` (llamax_b200.build.
install_reference). Used by `bench.py --impl reference` / `cpu_baseline` and tools/incumbents.py only — never by the
product path (llamax_b200/ has no import of this package).

The reference's `torchao::int8_mm_dequant` op has Meta and CUDA (Triton) implementations only
(subclasses/int8_mm.py:135-149); `load(cpu_shim=True)` registers the CPU key on the reference's own library object
with exact int32 accumulation (torch._int_mm) and the epilogue order of int8_mm.py:112-114, so that the
dynamic-INT8 mode of the reference runs on host cores. Nothing inside baseline/_ref is edited.
"""
import os
import sys

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_loaded = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "modelling")) and os.path.isdir(os.path.join(REF_DIR, "subclasses"))


def load(cpu_shim: bool = True):
    """Returns (modelling, subclasses) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference not installed under {REF_DIR}: run `python -m llamax_b200.build` where "
                           "/root/reference exists")
    if "mods" not in _loaded:
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import modelling  # noqa: F401  (reference)
        import subclasses  # noqa: F401  (reference)

        assert os.path.dirname(os.path.abspath(modelling.__file__)).startswith(REF_DIR), modelling.__file__
        _loaded["mods"] = (modelling, subclasses)
    if cpu_shim and "shim" not in _loaded:
        import torch
        from subclasses import int8_mm as ref_int8_mm

        @torch.library.impl(ref_int8_mm.lib, "int8_mm_dequant", "CPU")
        def _cpu_int8_mm_dequant(A, B, a_scale, b_scale):
            return (torch._int_mm(A, B).float() * a_scale.float().view(-1, 1) * b_scale.float().view(1, -1)).to(a_scale.dtype)

        _loaded["shim"] = _cpu_int8_mm_dequant
    return _loaded["mods"]
