"""CPU oracle for the batched eval-scoring forward (TEST INFRASTRUCTURE ONLY).

This package is a plain PyTorch fp32 restatement of the arithmetic on the
reference's scoring path (waveform -> XLS-R -> AASIST / Conformer -> logits).
It exists to *check* the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it; the
product package (``real-time-deepfake-speech-detection_b200``) never does and
fails loudly when its CUDA extension is missing.

Parity pinning status
---------------------
* ``aasist_ref.py`` / ``conformer_model_ref.py`` / ``preemph_ref.py`` restate
  files that live in ``/root/reference`` and are pinned **bit-exactly** against
  the reference's own ``models/*.py`` executed unmodified (over import shims)
  by ``oracle/check_against_reference.py``; golden vectors produced by that
  run are committed under ``tests/golden/``.
* ``wav2vec2_ref.py`` (fairseq ``Wav2Vec2Model``) and ``conformer_block_ref.py``
  (lucidrains ``conformer.ConformerBlock``) restate *third-party, un-vendored,
  un-pinned* dependencies that are absent from ``/root/reference`` and from
  this image.  The reference ships no tests or golden vectors for them, so
  for these two files **parity is unpinned by the reference itself**; they are
  cross-checked against two independent in-image witnesses
  (``torchaudio.models.wav2vec2_xlsr_300m`` and HF ``Wav2Vec2Model``) by
  ``oracle/check_against_witnesses.py``.
"""
