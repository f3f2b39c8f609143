"""CPU tests of the host logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, VoiceEncConfig, _lib, scheduler, synth
from oracle import frontend, nets, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cbx.h")).read()
    declared = set(re.findall(r"\b(cbx_[a-z_0-9]+)\s*\(", hdr))
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTS)


def test_num_wins_and_step_match_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "ints.npz"))
    for n, (w, t) in zip(g["n_frames"], g["wins77"]):
        assert _lib.num_wins(int(n), 77, 0.8) == (int(w), int(t))
    for n, (w, t) in zip(g["n_frames"], g["wins80"]):
        assert _lib.num_wins(int(n), 80, 0.8) == (int(w), int(t))
    assert _lib.frame_step(0.5, 1.3) == 77 and _lib.frame_step(0.5, None) == 80
    for rate in (0.7, 1.0, 1.3, 2.0, 3.3):
        assert _lib.frame_step(0.5, rate) == nets.frame_step(0.5, rate)
    for ov in (0.0, 0.25, 0.5, 0.9):
        assert _lib.frame_step(ov, None) == nets.frame_step(ov, None)


@pytest.mark.parametrize("n,want", [(160000, (1001, 12, 1007, 998, 499, 5)), (48000, (301, 3, 314, 298, 149, 2)),
                                     (480000, (3001, 38, 3009, 2998, 1499, 15)), (720, (5, 1, 160, 3, 2, 1)),
                                     (399, (3, 1, 160, 0, 0, 0))])
def test_plan_clip(n, want):
    p = _lib.plan_clip(n)
    assert (p.ve_frames, p.ve_partials, p.ve_target, p.xv_frames, p.xv_tdnn, p.xv_segments) == want


def test_plan_arithmetic_matches_oracle_on_random_inputs():
    """cbx_ve_num_wins / cbx_plan_clip against the oracle's restatement of get_num_wins (voice_encoder.py:54-66) and of the
    Kaldi frame count (kaldi.py:63-67) over random frame counts, steps, coverages and clip lengths (bit-exact integers)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=400, deadline=None)
    @given(st.integers(1, 200000), st.integers(1, 160), st.floats(0.0, 1.0))
    def wins(n_frames, step, cov):
        assert _lib.num_wins(n_frames, step, cov) == tuple(int(v) for v in nets.num_wins(n_frames, step, cov))

    @settings(max_examples=400, deadline=None)
    @given(st.integers(0, 16000 * 600), st.sampled_from([77, 80, 160, 1]))
    def plan(n, step):
        p = _lib.plan_clip(n, step, 0.8)
        w, t = nets.num_wins(1 + n // 160, step, 0.8)
        tk = frontend.kaldi_num_frames(n) if n >= 400 else 0
        td = (tk - 1) // 2 + 1 if tk > 0 else 0
        assert (p.ve_frames, p.ve_partials, p.ve_target, p.xv_frames, p.xv_tdnn, p.xv_segments) == (1 + n // 160, int(w), int(t), tk, td, -(-td // 100))

    wins()
    plan()


def test_no_cpu_fallback():
    with pytest.raises(_lib.CbxError):
        _lib.Context(0)
    ve = VoiceEncoder()
    with pytest.raises(_lib.CbxError):
        ve.embeds_from_wavs([synth.clip(0, 16000)], 16000)
    with pytest.raises(_lib.CbxError):
        CAMPPlus().inference([np.zeros(16000, np.float32)])


def test_state_dict_compat():
    ve = VoiceEncoder(VoiceEncConfig())
    ve.load_state_dict(weights.ve_state_dict("W0"), strict=True)
    cp = CAMPPlus()
    sd = weights.campplus_state_dict("W1")
    cp.load_state_dict(sd, strict=True)
    assert len(cp.state_dict()) == 937
    assert sum(v.numel() for k, v in cp.state_dict().items() if v.dtype.is_floating_point and "running" not in k) > 6.8e6


def test_baked_config_is_enforced():
    class HP(VoiceEncConfig):
        num_mels = 80
    with pytest.raises(ValueError):
        VoiceEncoder(HP())
    with pytest.raises(ValueError):
        CAMPPlus(embedding_size=256)


def test_scheduler_partition_balances_and_inverts():
    lens = synth.ragged_lengths(1024)
    for world in (1, 2, 4, 8):
        shards = scheduler.partition(lens, world)
        allidx = np.concatenate([s for s in shards])
        assert sorted(allidx.tolist()) == list(range(len(lens)))
        sizes = {len(s) for s in shards}
        assert max(sizes) - min(sizes) <= 1
        costs = [sum(_lib.clip_cost(int(lens[i])) for i in s) for s in shards]
        assert max(costs) / (sum(costs) / world) < 1.03
        inv = scheduler.inverse_permutation(shards, len(lens))
        gathered = np.concatenate([np.pad(s, (0, max(sizes) - len(s)), constant_values=-1) for s in shards])
        assert (gathered[inv] == np.arange(len(lens))).all()
        # the C entry point itself: rank, row inside the rank's shard and summed cost per rank
        rank_of, row_of, rank_cost = _lib.partition(lens, world)
        assert (inv == rank_of.astype(np.int64) * max(sizes) + row_of).all()
        np.testing.assert_allclose(rank_cost, costs, rtol=1e-12)
    for n in (0, 1, 3):                                   # fewer clips than ranks: empty shards, nothing lost
        shards = scheduler.partition(lens[:n], 4)
        assert sorted(np.concatenate(shards).tolist()) == list(range(n)) and max(len(s) for s in shards) <= 1
    with pytest.raises(_lib.CbxError):
        _lib.partition([16000, -1], 2)
    with pytest.raises(_lib.CbxError):
        _lib.partition([16000], 0)


def test_voice_profile_container_round_trip(tmp_path):
    """The .npy pickle-dict container of s3gen.py:427-470 / tts.py:537-586: keys, shapes, dtypes, optional fields, old-format files."""
    import torch
    from chatterbox_embed_b200 import voice_profile as vp
    rng = np.random.RandomState(0)
    emb, ve = rng.randn(1, 192).astype(np.float32), rng.randn(1, 256).astype(np.float32)
    feat = rng.randn(1, 37, 80).astype(np.float32)
    tok, tok_len = rng.randint(0, 6561, (1, 18)).astype(np.int64), np.array([18], np.int64)
    prof = vp.VoiceProfile(torch.from_numpy(emb), torch.from_numpy(feat), None, torch.from_numpy(tok), torch.from_numpy(tok_len))
    p = str(tmp_path / "profile.npy")
    np.save(p, vp.profile_dict(prof, ve_embedding=torch.from_numpy(ve)))
    raw = np.load(p, allow_pickle=True).item()                     # what the reference's readers do (tts.py:559, s3gen.py:449)
    assert list(raw) == ["embedding", "ve_embedding", "prompt_feat", "prompt_token", "prompt_token_len"]
    assert raw["embedding"].dtype == np.float32 and raw["embedding"].shape == (1, 192) and raw["prompt_token"].dtype == np.int64
    back = vp.load_voice_profile(p)
    assert torch.equal(back.embedding, torch.from_numpy(emb)) and torch.equal(back.ve_embedding, torch.from_numpy(ve))
    assert torch.equal(back.prompt_feat, torch.from_numpy(feat)) and back.prompt_feat_len is None
    assert torch.equal(back.prompt_token, torch.from_numpy(tok)) and torch.equal(back.prompt_token_len, torch.from_numpy(tok_len))
    # VoiceProfile.save / .load (no ve_embedding key: the "old format" branch of vc.py:697-700)
    p2 = str(tmp_path / "plain.npy")
    vp.VoiceProfile(torch.from_numpy(emb)).save(p2)
    assert list(np.load(p2, allow_pickle=True).item()) == ["embedding"]
    old = vp.load_voice_profile(p2)
    assert old.ve_embedding is None and old.prompt_feat is None and old.prompt_token is None
    assert torch.equal(vp.VoiceProfile.load(p2).embedding, torch.from_numpy(emb))


def test_load_audio_like_librosa_load(tmp_path):
    from scipy.io import wavfile
    from chatterbox_embed_b200 import voice_profile as vp
    rng = np.random.RandomState(1)
    pcm = rng.randint(-32768, 32767, (1000, 2)).astype(np.int16)
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 22050, pcm)
    y, sr = vp.load_audio(p)
    assert sr == 22050 and y.dtype == np.float32 and y.shape == (1000,)
    assert np.array_equal(y, (pcm.astype(np.float32) / 32768.0).mean(axis=1))         # soundfile scaling + librosa to_mono


def test_s3_frontend_pad_matches_reference_fixture(golden_dir):
    """S3Tokenizer.pad (s3tokenizer.py:52-74) is host arithmetic: lengths from the verbatim reference."""
    import torch
    from chatterbox_embed_b200.s3tokenizer import S3TokenizerFrontend
    g = np.load(os.path.join(golden_dir, "ref_s3_log_mel.npz"))
    fe = S3TokenizerFrontend("cpu")                      # pad / _prepare_audio never touch the device
    got = [fe.pad([np.zeros(int(n), np.float32)], 16000)[0].shape[1] for n in g["pad_in"]]
    assert got == g["pad_out"].tolist()
    out = fe.pad([torch.ones(2, 700)], 16000)[0]         # (B, L) tensors are padded on the right with zeros
    assert tuple(out.shape) == (2, 1280) and float(out[:, 700:].abs().max()) == 0.0
    assert [tuple(w.shape) for w in fe._prepare_audio([np.zeros(5, np.float32), torch.zeros(1, 7)])] == [(1, 5), (1, 7)]


def test_next_row_wrappers_refuse_what_they_do_not_build():
    """No silent fallbacks: non-default mel settings, a tokenizer forward without a quantizer, CPU tensors."""
    import torch
    from chatterbox_embed_b200 import _lib, mel_spectrogram, Resample
    from chatterbox_embed_b200.s3tokenizer import S3TokenizerFrontend
    with pytest.raises(NotImplementedError):
        mel_spectrogram(torch.zeros(1, 24000), n_fft=1024)
    with pytest.raises(NotImplementedError):
        S3TokenizerFrontend("cpu").forward([np.zeros(16000, np.float32)])
    with pytest.raises(Exception, match="integer"):
        Resample(44100.5, 16000)
    assert Resample(16000, 16000)(torch.ones(3)) is not None         # identity needs no device
    assert _lib.prompt_mel_frames(240000) == 500 and _lib.s3_log_mel_frames(160000) == 1000 and _lib.resample_out_len(44100, 16000, 441000) == 160000
    if not torch.cuda.is_available():
        with pytest.raises(_lib.CbxError):
            mel_spectrogram(torch.zeros(1, 24000))                    # needs a B200: no CPU path
