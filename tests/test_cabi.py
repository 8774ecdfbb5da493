"""CPU: the C-ABI library builds, loads, and exports every symbol include/llamax_b200.h declares (no compute)."""
import ctypes
import os
import re

from llamax_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "llamax_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(?:int|const char\*)\s+(llamax_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    return {name: [a for a in args.split(",") if a.strip() not in ("", "void")] for name, args in decls}


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 18
    for name, args in declared.items():
        fn = getattr(lib, name)  # raises AttributeError when missing
        assert fn is not None
        if name in _lib.SIGNATURES:
            assert len(_lib.SIGNATURES[name]) == len(args), f"{name}: binding has {len(_lib.SIGNATURES[name])} args, header {len(args)}"
    assert set(_lib.SIGNATURES) <= set(declared)
    assert lib.llamax_version() == 102
    assert isinstance(lib.llamax_last_error(), bytes)


def test_epilogue_struct_layout_matches_header():
    e = _lib.Epilogue
    assert [f[0] for f in e._fields_] == ["lora_h", "ldh", "lora_b", "lora_rank", "lora_scale", "resid", "ldr", "seg_n0",
                                          "seg_n1", "rope", "rope_S", "rope_cols"]
    assert ctypes.sizeof(e) == 72 and e.lora_rank.offset == 24 and e.lora_scale.offset == 28 and e.resid.offset == 32
    assert e.seg_n0.offset == 48 and e.seg_n1.offset == 52
    assert e.rope.offset == 56 and e.rope_S.offset == 64 and e.rope_cols.offset == 68      # ABI version 102


def test_copy_job_struct_layout_matches_header():
    j = _lib.CopyJob
    assert [f[0] for f in j._fields_] == ["src", "dst", "src_ld", "dst_ld", "rows", "cols", "scale", "flags"]
    assert ctypes.sizeof(j) == 48 and j.rows.offset == 32 and j.scale.offset == 40 and j.flags.offset == 44


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    lib = _lib.load()
    rc = lib.llamax_rowquant_int8(None, 0, None, None, 4, 16, None)
    assert rc == -1 and b"null" in lib.llamax_last_error()
    rc = lib.llamax_set_gemm_cta_group(3)
    assert rc == -1
    assert lib.llamax_set_gemm_cta_group(2) == 0
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.llamax_attn_fwd(p, 64, p, 64, p, 64, p, 64, p, 1, 16, 2, 1, 32, 0, None, None, 1.0, None)
    assert rc == -1 and b"head_dim" in lib.llamax_last_error()


def test_new_gemm_entry_points_validate_arguments_without_gpu():
    """bf16_gemm_tn / bf16_gemm_swiglu_bwd reject null pointers and unsupported shapes before any launch."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.llamax_bf16_gemm_tn(None, 8, p, 8, p, 8, 8, 8, 8, None) == -1
    assert b"null" in lib.llamax_last_error()
    assert lib.llamax_bf16_gemm_tn(p, 8, p, 8, p, 8, 12, 8, 8, None) == -1          # M % 8 != 0
    assert b"multiple of 8" in lib.llamax_last_error()
    # fused SwiGLU-backward epilogue: N % 16, pitches >= 2N, 16-byte alignment
    assert lib.llamax_bf16_gemm_swiglu_bwd(p, 64, p, 64, 16, 24, 64, None, p, 48, p, 48, None, None) == -1
    assert b"N % 16" in lib.llamax_last_error()
    assert lib.llamax_bf16_gemm_swiglu_bwd(p, 64, p, 64, 16, 32, 64, None, p, 48, p, 64, None, None) == -1   # ld_ab < 2N
    assert lib.llamax_bf16_gemm_swiglu_bwd(p, 64, p, 64, 16, 32, 64, None, None, 64, p, 64, None, None) == -1
    assert b"null" in lib.llamax_last_error()
    mis = ctypes.c_void_p(p.value + 2)
    assert lib.llamax_bf16_gemm_swiglu_bwd(p, 64, p, 64, 16, 32, 64, None, mis, 64, p, 64, None, None) == -1
    assert b"aligned" in lib.llamax_last_error()


def test_round2_entry_points_validate_arguments_without_gpu():
    """Mixed-input GEMMs, the row-dot GEMM and the segmented LoRA epilogue reject bad arguments before any launch."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(512)
    p = ctypes.cast(buf, ctypes.c_void_p)
    # mixed-input: null pointers; layout 0 has no tail; layout must be 0 / 1; K1 multiple of 64; tail and K - K1 agree
    assert lib.llamax_bf16_int8_gemm(None, 64, p, 64, p, 0, None, 0, p, 64, 16, 64, 64, 64, None, None) == -1
    assert b"null" in lib.llamax_last_error()
    assert lib.llamax_bf16_int8_gemm(p, 64, p, 64, p, 0, p, 64, p, 64, 16, 64, 64, 64, None, None) == -1
    assert b"no tail" in lib.llamax_last_error()
    assert lib.llamax_bf16_int8_gemm(p, 64, p, 64, p, 2, None, 0, p, 64, 16, 64, 64, 64, None, None) == -1
    assert b"b_layout" in lib.llamax_last_error()
    assert lib.llamax_bf16_int8_gemm(p, 96, p, 64, p, 1, None, 0, p, 64, 16, 64, 96, 96, None, None) == -1
    assert b"multiple of 64" in lib.llamax_last_error()
    assert lib.llamax_bf16_int8_gemm(p, 128, p, 64, p, 1, None, 0, p, 64, 16, 64, 72, 64, None, None) == -1
    assert b"disagree" in lib.llamax_last_error()
    assert lib.llamax_bf16_int8_gemm_swiglu_bwd(p, 64, p, 64, None, 16, 32, 64, None, p, 64, p, 64, None, None) == -1
    assert b"null" in lib.llamax_last_error()
    # row-dot GEMM: N % 256, M % S, aliasing
    assert lib.llamax_bf16_gemm_rowdot(p, 64, p, 64, p, 128, 16, 128, 64, None, p, 128, p, 16, None) == -1
    assert b"multiple of 256" in lib.llamax_last_error()
    assert lib.llamax_bf16_gemm_rowdot(p, 64, p, 64, p, 256, 16, 256, 64, None, p, 256, p, 5, None) == -1
    assert b"multiple of S" in lib.llamax_last_error()
    q = ctypes.cast(ctypes.create_string_buffer(64), ctypes.c_void_p)
    assert lib.llamax_bf16_gemm_rowdot(p, 64, p, 64, q, 256, 16, 256, 64, None, q, 256, p, 16, None) == -1
    assert b"alias" in lib.llamax_last_error()
    # LoRA column segments must be increasing multiples of the tile width
    ep = _lib.Epilogue()
    ep.lora_h, ep.ldh, ep.lora_b, ep.lora_rank, ep.lora_scale = p.value, 24, p.value, 8, 1.0
    ep.seg_n0, ep.seg_n1 = 100, 200
    assert lib.llamax_int8_gemm_dequant(p, 64, p, 64, p, p, p, 768, 16, 768, 64, ctypes.byref(ep), None) == -1
    assert b"segments" in lib.llamax_last_error()
    # attention backward accepts o = NULL (delta given) but still rejects the other null pointers
    assert lib.llamax_attn_bwd(p, 64, p, 64, p, 64, None, 0, p, None, 64, p, 64, p, 64, p, 64, p, p,
                               1, 16, 2, 1, 128, 0, None, None, None, 1.0, None, None) == -1
    assert b"null" in lib.llamax_last_error()
