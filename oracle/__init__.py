"""CPU oracle for the Whisper-Aries hot path (log-mel -> Whisper encoder).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
thing timed on the host cores beside the GPU number).  The product package
``whisper_aries_b200`` never imports this package and fails loudly when its
CUDA library is missing.

PARITY UNPINNED against the true reference: the arithmetic of the path lives
in un-vendored wheels (faster-whisper==1.1.1, ctranslate2==4.6.0, numpy==1.26.4;
ref: requirements.txt:12,9,36) that are not installable in this sandbox and the
reference repo holds no golden vector for log-mel values or encoder hidden
states.  The restatement is therefore anchored on
  * the published algorithm of those pinned versions (restated, not copied),
  * the reference's call sites (final_optimized_transcriber.py:326,189;
    conversation_transcriber.py:72-77), and
  * an independent implementation that IS importable here (HF transformers
    5.5 WhisperFeatureExtractor / WhisperEncoder), cross-checked by
    ``oracle/make_golden.py`` and ``tests/test_oracle_*.py``.
"""
