"""B200-native speaker-embedding path of Chatterbox (VoiceEncoder + CAMPPlus) behind the reference's Python API."""
from .config import VoiceEncConfig
from .voice_encoder import VoiceEncoder, get_frame_step, get_num_wins
from .campplus import CAMPPlus
from .s3gen_cond import SpeakerConditioner
from .resample import Resample, get_resampler
from .mel import mel_spectrogram
from .consumers import SpeakerProjections
from .voice_profile import VoiceProfile, VoiceProfiler, load_voice_profile

__all__ = ["VoiceEncConfig", "VoiceEncoder", "CAMPPlus", "SpeakerConditioner", "Resample", "get_resampler", "mel_spectrogram", "SpeakerProjections", "VoiceProfile", "VoiceProfiler", "load_voice_profile", "get_frame_step", "get_num_wins"]
