"""GPU parity of the audio stem: conv-as-GEMM over overlapping-row views + fused bias/GELU
passes + col2im, against the fp32 evaluation of the reference module sequence (modelling/audio.py:26-31)."""
import pytest
import torch

from llamax_b200.modelling.audio import AudioStemFn
from oracle import ref_ops as R
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,Cin,T,C", [(2, 80, 200, 512), (3, 128, 50, 256), (1, 80, 1024, 512)])
def test_audio_stem_fwd_bwd_vs_oracle(B, Cin, T, C):
    torch.manual_seed(B * T + C)
    conv1 = torch.nn.Conv1d(Cin, C, 3, 1, 1)
    conv2 = torch.nn.Conv1d(C, C, 3, 2, 1)
    mel = torch.randn(B, Cin, T).bfloat16()
    dout = torch.randn(B, T // 2, C).bfloat16()
    params = [p.detach().bfloat16() for p in (conv1.weight, conv1.bias, conv2.weight, conv2.bias)]
    # oracle: same bf16-rounded inputs, fp32 arithmetic
    ref_in = [p.float().requires_grad_(True) for p in params]
    y_ref = R.audio_stem_ref(mel.float(), *ref_in)
    y_ref.backward(dout.float())
    ours_in = [p.cuda().requires_grad_(True) for p in params]
    y = AudioStemFn.apply(mel.cuda(), *ours_in)
    assert y.shape == (B, T // 2, C)
    assert rel_err(y, y_ref) <= 1e-2
    y.backward(dout.cuda())
    for name, a, b in zip(("dw1", "db1", "dw2", "db2"), ours_in, ref_in):
        assert a.grad.shape == b.grad.shape
        assert rel_err(a.grad, b.grad) <= 1e-2, name


def test_llama_audio_uses_own_stem_and_matches_module_path():
    """The model-level switch: LlamaAudio.embed_audio through AudioStemFn vs through nn.Conv1d / nn.GELU on the same
    parameters (both bf16 on the GPU)."""
    from tests.helpers import build_tiny_llama

    model = build_tiny_llama(True, num_layers=1, audio=True).cuda()
    model.build_cache()
    torch.manual_seed(1)
    audio = torch.randn(2, 16000, device="cuda")
    import llamax_b200.modelling.audio as A

    emb_own = model.embed_audio(audio)
    A._OWN_STEM = False
    try:
        emb_lib = model.embed_audio(audio)
    finally:
        A._OWN_STEM = True
    assert emb_own.shape == emb_lib.shape
    assert rel_err(emb_own, emb_lib) <= 2e-2
