"""Hyper-parameters of the VoiceEncoder path (mirror of voice_encoder/config.py:1-18).  The CUDA kernels are
compiled for exactly these values; constructing a VoiceEncoder with other values raises."""


class VoiceEncConfig:
    num_mels = 40
    sample_rate = 16000
    speaker_embed_size = 256
    ve_hidden_size = 256
    flatten_lstm_params = False
    n_fft = 400
    hop_size = 160
    win_size = 400
    fmax = 8000
    fmin = 0
    preemphasis = 0.
    mel_power = 2.0
    mel_type = "amp"
    normalized_mels = False
    ve_partial_frames = 160
    ve_final_relu = True
    stft_magnitude_min = 1e-4


_BAKED = dict(num_mels=40, sample_rate=16000, speaker_embed_size=256, ve_hidden_size=256, n_fft=400, hop_size=160,
              win_size=400, fmax=8000, fmin=0, preemphasis=0., mel_power=2.0, mel_type="amp", normalized_mels=False,
              ve_partial_frames=160, ve_final_relu=True)


def check_baked(hp) -> None:
    for k, v in _BAKED.items():
        if getattr(hp, k) != v:
            raise ValueError(f"VoiceEncConfig.{k}={getattr(hp, k)!r} differs from the value the sm_100a kernels are built for ({v!r})")
