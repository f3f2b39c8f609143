"""CPU oracle for the voice-clone conditioning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``chatterbox_embed_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker
or as the timed CPU baseline -- never as the thing shipped.

It is a restatement (numpy + torch-CPU functional ops) of what the reference
computes on this path:

* ``frontend.py``  -- librosa 0.11 semantics used by ``melspec.py:9-64`` and
  ``voice_encoder.py:267`` (stft / filters.mel / effects.trim), and the Kaldi
  fbank of ``xvector.py:45-58`` (torchaudio.compliance.kaldi).
* ``nets.py``      -- ``VoiceEncoder.forward/inference`` (``voice_encoder.py:139-199``)
  and ``CAMPPlus.forward`` (``xvector.py:61-423``) written against a plain
  ``state_dict`` with the reference's key names.
* ``weights.py``   -- seeded weight sets W0/W1/W2.
* ``synth.py``     -- synthetic audio (white noise / chirps).
* ``refload.py``   -- loads the *verbatim* reference modules from /root/reference
  (only possible in the build container; used by ``make_golden.py`` and by the
  tests that are skipped when the reference tree is absent).

Parity pinning: the CAMPPlus path and the VoiceEncoder mel->embedding path are
pinned against the verbatim reference modules (``tests/golden/*.npz`` were
produced by ``oracle/make_golden.py`` running the reference itself).  The
VoiceEncoder *front-end* (librosa stft/mel/trim) is **parity unpinned at the
librosa boundary**: librosa 0.11.0 is neither vendored in the reference nor
installed/installable here, and the reference holds no golden vectors for it.
It is cross-checked against ``torch.stft`` and
``torchaudio.functional.melscale_fbanks`` instead (tests/test_oracle.py).
"""
