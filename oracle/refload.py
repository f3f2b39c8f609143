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


def make_voice_encoder(sd):
    ve = voice_encoder_module().VoiceEncoder()
    ve.load_state_dict(sd, strict=True)
    return ve.eval()


def make_campplus(sd):
    m = xvector_module().CAMPPlus()
    m.load_state_dict(sd, strict=True)
    return m.eval()
