"""Llama with an audio prefix — drop-in for modelling/audio.py of the reference.

Front-end (mel-spectrogram -> log10 -> cepstral mean normalisation -> Conv1d(k3,s1)+GELU -> Conv1d(k3,s2)+GELU,
audio.py:26-31,51-60) is unchanged PyTorch: it is a few percent of the FLOPs and listed as a later scope row.
The decoder blocks behind it run on the fused sm_100a path.

Extension over the reference: `LlamaAudio.forward(..., prefix_lm=True)` makes the audio positions a bidirectional
prefix (mask(q, kv) = (kv < P) | (q >= kv), P = number of audio positions) — the prefix-LM objective the
reference's README plans (README.md:16) but runs as plain causal attention (audio.py:65-70). The default
(`prefix_lm=False`) keeps the reference behaviour.
"""

from typing import NamedTuple

import torch
from torch import Tensor, nn

from .llama import Llama, LlamaConfig, PrefixLM


class AudioConfig(NamedTuple):
    sample_rate: int = 16_000
    n_fft: int = 512
    win_length: int = 400
    hop_length: int = 160
    n_mels: int = 128


class LlamaAudio(Llama):
    def __init__(self, config: LlamaConfig, audio_config: AudioConfig = AudioConfig()):
        super().__init__(config)
        self.audio_config = audio_config
        # Whisper-style stem: stride-1 conv then stride-2 conv, GELU after each
        self.audio_embed = nn.Sequential(
            nn.Conv1d(audio_config.n_mels, config.embed_dim, 3, 1, 1),
            nn.GELU(),
            nn.Conv1d(config.embed_dim, config.embed_dim, 3, 2, 1),
            nn.GELU(),
        )

    def build_cache(self, inference: bool = False):
        super().build_cache(inference)
        from torchaudio.transforms import MelSpectrogram

        self.melspec = MelSpectrogram(**self.audio_config._asdict(), norm="slaney", mel_scale="slaney")
        self.melspec.to(self.tok_embeddings.weight.device)

    def embed_audio(self, audio: Tensor) -> Tensor:
        """waveform [B, T] -> prefix embeddings [B, P, embed_dim]."""
        mel = self.melspec(audio)[..., :-1].clip(1e-12).log10()  # drop the last frame: even length
        mel = mel - mel.mean(2, keepdim=True)
        mel = mel.to(dtype=self.tok_embeddings.weight.dtype)
        if self.config.activation_checkpointing:
            from torch.utils.checkpoint import checkpoint

            emb = checkpoint(self.audio_embed, mel, use_reentrant=False)
        else:
            emb = self.audio_embed(mel)
        return emb.transpose(1, 2)

    def forward(self, audio: Tensor | None, tokens: Tensor, *, input_pos: Tensor | None = None,
                labels: Tensor | None = None, prefix_lm: bool = False) -> Tensor:
        if input_pos is not None:
            raise NotImplementedError("llamax_b200: input_pos (inference) is outside the fine-tuning hot path")
        x = self.tok_embeddings(tokens)
        n_prefix = 0
        if audio is not None:
            prefix = self.embed_audio(audio)
            n_prefix = prefix.shape[1]
            x = torch.cat([prefix, x], dim=1)
        block_mask = PrefixLM(n_prefix) if (prefix_lm and n_prefix > 0) else None
        x = self._run_layers(x, block_mask)
        if n_prefix:
            x = x[:, n_prefix:]  # loss / logits on text positions only
        return self._head(x, labels)
