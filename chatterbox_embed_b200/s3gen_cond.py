"""The voice-clone ``.npy`` boundary of S3Token2Mel (s3gen.py:107-119, 145-148) and ChatterboxTTS.save_voice_clone
(tts.py:502-508), for 16 kHz input: CAMPPlus x-vector -> ``np.save`` of a (1,192) float32 array (NPY v1, C order),
byte-compatible with ``audio_test/reference_voice_clone.npy``."""
from __future__ import annotations

from typing import Union

import numpy as np
import torch

from .campplus import CAMPPlus
from .resample import get_resampler

S3_SR = 16000


class SpeakerConditioner:
    """Holds the speaker encoder of S3Token2Mel and reproduces its clone save/load methods."""

    def __init__(self, speaker_encoder: CAMPPlus):
        self.speaker_encoder = speaker_encoder

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
