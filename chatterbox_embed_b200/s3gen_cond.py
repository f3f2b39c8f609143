"""The speaker-conditioning methods of S3Token2Mel: the voice-clone ``.npy`` boundary (s3gen.py:107-119, 145-148;
ChatterboxTTS.save_voice_clone tts.py:502-508) -- CAMPPlus x-vector -> ``np.save`` of a (1,192) float32 array (NPY v1, C
order), byte-compatible with ``audio_test/reference_voice_clone.npy`` -- and ``embed_ref`` (s3gen.py:150-207): resample to
24 kHz / 16 kHz, prompt mel, x-vector.  The S3 tokenizer is a separate model outside this path; pass one in to get tokens."""
from __future__ import annotations

from typing import Union

import numpy as np
import torch

from .campplus import CAMPPlus
from .mel import mel_spectrogram
from .resample import get_resampler

S3_SR = 16000
S3GEN_SR = 24000


class SpeakerConditioner:
    """Holds the speaker encoder of S3Token2Mel and reproduces its clone save/load methods."""

    def __init__(self, speaker_encoder: CAMPPlus, tokenizer=None):
        self.speaker_encoder = speaker_encoder
        self.mel_extractor = mel_spectrogram
        self.tokenizer = tokenizer          # S3Tokenizer-like callable (wav_16k) -> (tokens, lens); not part of this package

    @property
    def device(self):
        return self.speaker_encoder.device

    @torch.inference_mode()
    def save_voice_clone(self, ref_wav: Union[torch.Tensor, np.ndarray], ref_sr: int, save_path: str):
        if isinstance(ref_wav, np.ndarray):
            ref_wav = torch.from_numpy(ref_wav).float()
        if len(ref_wav.shape) == 1:
            ref_wav = ref_wav.unsqueeze(0)
        ref_wav = ref_wav.to(self.device)
        ref_wav_16 = get_resampler(ref_sr, S3_SR, self.device)(ref_wav)        # s3gen.py:116 (identity at 16 kHz)
        embedding = self.speaker_encoder.inference(ref_wav_16)
        np.save(save_path, embedding.detach().cpu().numpy())

    @torch.inference_mode()
    def load_voice_clone(self, embedding_path: str) -> torch.Tensor:
        emb = np.load(embedding_path)
        return torch.from_numpy(emb).to(self.device)

    @torch.inference_mode()
    def embed_ref(self, ref_wav: Union[torch.Tensor, np.ndarray], ref_sr: int, device="auto", ref_fade_out=True):
        """s3gen.py:150-207.  ``prompt_token`` / ``prompt_token_len`` are None unless a tokenizer was given."""
        device = self.device if device == "auto" else device
        if isinstance(ref_wav, np.ndarray):
            ref_wav = torch.from_numpy(ref_wav).float()
        if ref_wav.device != device:
            ref_wav = ref_wav.to(device)
        if len(ref_wav.shape) == 1:
            ref_wav = ref_wav.unsqueeze(0)
        if ref_wav.size(1) > 10 * ref_sr:
            print("WARNING: cosydec received ref longer than 10s")
        ref_wav_24 = ref_wav
        if ref_sr != S3GEN_SR:
            ref_wav_24 = get_resampler(ref_sr, S3GEN_SR, device)(ref_wav)
        ref_mels_24 = self.mel_extractor(ref_wav_24).transpose(1, 2).to(device)
        ref_wav_16 = get_resampler(ref_sr, S3_SR, device)(ref_wav).to(device)
        ref_x_vector = self.speaker_encoder.inference(ref_wav_16)
        tokens = token_lens = None
        if self.tokenizer is not None:
            tokens, token_lens = self.tokenizer(ref_wav_16)
            if ref_mels_24.shape[1] != 2 * tokens.shape[1]:
                tokens = tokens[:, :ref_mels_24.shape[1] // 2]
                token_lens = token_lens.clone().detach()
                token_lens[0] = tokens.shape[1]
            tokens = tokens.to(device)
        return dict(prompt_token=tokens, prompt_token_len=token_lens, prompt_feat=ref_mels_24, prompt_feat_len=None,
                    embedding=ref_x_vector)
