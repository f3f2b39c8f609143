"""B200-native speaker-embedding path of Chatterbox (VoiceEncoder + CAMPPlus) behind the reference's Python API."""
from .config import VoiceEncConfig
from .voice_encoder import VoiceEncoder, get_frame_step, get_num_wins
from .campplus import CAMPPlus
from .s3gen_cond import SpeakerConditioner
from .resample import Resample, get_resampler
from .mel import mel_spectrogram

__all__ = ["VoiceEncConfig", "VoiceEncoder", "CAMPPlus", "SpeakerConditioner", "Resample", "get_resampler", "mel_spectrogram", "get_frame_step", "get_num_wins"]
