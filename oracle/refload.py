"""Oracle (TEST INFRASTRUCTURE): import the VERBATIM reference modules from
/root/reference (build container only -- the tree does not exist on the GPU box).
Recipe of SURVEY.md Appendix C: ``xvector.py`` is loaded by path; the
``voice_encoder`` directory is mounted as a synthetic package with the oracle's
librosa shim registered as ``librosa``.  Nothing is copied into the repo.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

from . import frontend

REF_ROOT = os.environ.get("CBX_REFERENCE_ROOT", "/root/reference")
_MODELS = os.path.join(REF_ROOT, "src", "chatterbox", "models")


def available() -> bool:
    return os.path.isfile(os.path.join(_MODELS, "s3gen", "xvector.py"))


def xvector_module():
    if "ref_xvector" in sys.modules:
        return sys.modules["ref_xvector"]
    spec = importlib.util.spec_from_file_location("ref_xvector", os.path.join(_MODELS, "s3gen", "xvector.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_xvector"] = mod
    spec.loader.exec_module(mod)
    return mod


def voice_encoder_module():
    if "ref_ve_pkg.voice_encoder" in sys.modules:
        return sys.modules["ref_ve_pkg.voice_encoder"]
    frontend.register_librosa_shim()
    pkg = types.ModuleType("ref_ve_pkg")
    pkg.__path__ = [os.path.join(_MODELS, "voice_encoder")]
    sys.modules["ref_ve_pkg"] = pkg
    return importlib.import_module("ref_ve_pkg.voice_encoder")


def mel_module():
    """s3gen/utils/mel.py verbatim (its ``from librosa.filters import mel`` resolves to the oracle shim)."""
    if "ref_s3gen_mel" in sys.modules:
        return sys.modules["ref_s3gen_mel"]
    frontend.register_librosa_shim()
    spec = importlib.util.spec_from_file_location("ref_s3gen_mel", os.path.join(_MODELS, "s3gen", "utils", "mel.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_s3gen_mel"] = mod
    spec.loader.exec_module(mod)
    return mod


def s3tokenizer_module():
    """s3tokenizer/s3tokenizer.py verbatim.  Its base class comes from the third-party ``s3tokenizer`` package (pip, unpinned
    in the reference's pyproject, absent here): a stub with the three names the file imports stands in, so that the
    reference's own ``pad`` / ``log_mel_spectrogram`` run unmodified.  ``quantize`` (the network) is not available."""
    if "ref_s3tokenizer" in sys.modules:
        return sys.modules["ref_s3tokenizer"]
    import torch
    frontend.register_librosa_shim()

    class ModelConfig:
        n_mels = 128

    class S3TokenizerV2(torch.nn.Module):
        def __init__(self, name, config=ModelConfig()):
            super().__init__()

        @property
        def device(self):
            return torch.device("cpu")

    stub = types.ModuleType("s3tokenizer")
    stub.utils = types.ModuleType("s3tokenizer.utils")
    stub.utils.padding = None
    stub.model_v2 = types.ModuleType("s3tokenizer.model_v2")
    stub.model_v2.S3TokenizerV2, stub.model_v2.ModelConfig = S3TokenizerV2, ModelConfig
    for k, v in (("s3tokenizer", stub), ("s3tokenizer.utils", stub.utils), ("s3tokenizer.model_v2", stub.model_v2)):
        sys.modules.setdefault(k, v)
    spec = importlib.util.spec_from_file_location("ref_s3tokenizer", os.path.join(_MODELS, "s3tokenizer", "s3tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_s3tokenizer"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_voice_encoder(sd):
    ve = voice_encoder_module().VoiceEncoder()
    ve.load_state_dict(sd, strict=True)
    return ve.eval()


def make_campplus(sd):
    m = xvector_module().CAMPPlus()
    m.load_state_dict(sd, strict=True)
    return m.eval()
