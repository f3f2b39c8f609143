"""Hyper-parameters of the VoiceEncoder path: the same attribute names and values as the reference's ``VoiceEncConfig``
(voice_encoder/config.py:1-18).  The sm_100a kernels are compiled for exactly the values in ``_BAKED``; constructing a
VoiceEncoder with other values raises."""

# What the kernels are built for, grouped by the stage that consumes it.
_BAKED = dict(
    # front end (ve.cu / frontend_tc.cu): 16 kHz, periodic Hann 400, hop 160, 40 Slaney mels over 0..8000 Hz, power spectrum
    sample_rate=16000, n_fft=400, win_size=400, hop_size=160, num_mels=40, fmin=0, fmax=8000,
    preemphasis=0., mel_power=2.0, mel_type="amp", normalized_mels=False,
    # recurrence (lstm_tc.cu) and projection (project.cu): 160-frame partials, 3 x 256 LSTM, 256-d embedding behind a ReLU
    ve_partial_frames=160, ve_hidden_size=256, speaker_embed_size=256, ve_final_relu=True)

# Attributes the reference class also carries and this path never reads (a torch LSTM detail; the floor of the dB mel type).
_UNUSED = dict(flatten_lstm_params=False, stft_magnitude_min=1e-4)

VoiceEncConfig = type("VoiceEncConfig", (), {**_BAKED, **_UNUSED, "__doc__": "Class attributes as in voice_encoder/config.py:1-18."})


def check_baked(hp) -> None:
    for k, v in _BAKED.items():
        if getattr(hp, k) != v:
            raise ValueError(f"VoiceEncConfig.{k}={getattr(hp, k)!r} differs from the value the sm_100a kernels are built for ({v!r})")
