"""Row a-0: INTEGRATION.md section 2's patch, EXECUTED -- against a stand-in for ``faster_whisper.WhisperModel``.

faster-whisper / ctranslate2 are not installable offline, so the real class cannot be patched here.  ``UpstreamLikeModel``
below reproduces only the control flow of upstream 1.1.1 ``transcribe`` / ``generate_segments`` around the two hot-path
call sites (``self.feature_extractor(audio, chunk_length=...)`` and ``self.encode(segment)``, [unverified offline]) and
reads exactly the attributes upstream reads from the extractor.  The test applies the two assignments of INTEGRATION.md
to it and checks that the patched object (a) runs, (b) feeds ``encode`` what the oracle says upstream feeds it and (c)
returns encoder states matching the oracle -- i.e. that the shims really are drop-ins for those two attributes."""
import numpy as np
import pytest

from oracle import encoder as oenc, logmel as omel, synth as osynth

pytestmark = pytest.mark.gpu


class UpstreamLikeModel:
    """The slice of faster_whisper.WhisperModel that ``transcribe`` exercises on the hot path."""

    def __init__(self, feature_extractor, encode):
        self.feature_extractor = feature_extractor          # upstream: FeatureExtractor(**feat_kwargs)
        self.encode = encode                                # upstream: method wrapping ctranslate2 Whisper.encode
        self.seen_segments = []

    def transcribe(self, audio, chunk_length=None):
        fx = self.feature_extractor
        # attributes upstream reads from the extractor (transcribe.py): all must exist on the replacement
        for attr in ("n_fft", "hop_length", "chunk_length", "n_samples", "nb_max_frames", "time_per_frame",
                     "sampling_rate", "mel_filters"):
            assert hasattr(fx, attr), attr
        features = fx(audio, chunk_length=chunk_length)                      # [n_mels, (N + 160) // 160]
        content_frames = features.shape[-1] - 1
        seek, outs = 0, []
        while seek < content_frames:
            segment_size = min(fx.nb_max_frames, content_frames - seek)
            segment = features[:, seek: seek + segment_size]
            segment = np.pad(segment, [(0, 0), (0, fx.nb_max_frames - segment.shape[-1])])   # pad_or_trim
            self.seen_segments.append(segment)
            outs.append(self.encode(segment))
            seek += segment_size                                               # (no timestamp tokens: full windows)
        return outs


def test_integration_md_patch_runs_and_matches_the_oracle():
    import torch
    import whisper_aries_b200 as aries
    from whisper_aries_b200 import synthetic
    shape = synthetic.SHAPES["micro"]
    weights = synthetic.encoder_weights(shape, 1234)
    # --- INTEGRATION.md section 2, verbatim modulo the names of the stand-in
    model = UpstreamLikeModel(feature_extractor=None, encode=None)
    model.feature_extractor = aries.FeatureExtractor(feature_size=shape.n_mels, device="cuda:0")
    b200 = aries.WhisperEncoder(shape, weights, device="cuda:0")
    model.encode = lambda features: b200.encode(features)     # upstream wraps this in ctranslate2.StorageView.from_array
    # --- a 65-s call, as the reference's 185-s work items: three windows, the last one ragged
    audio = np.concatenate([osynth.window_signal(70), 0.3 * osynth.window_signal(71), osynth.window_signal(72)[:80000]])
    outs = model.transcribe(audio)
    assert len(outs) == 3 and all(o.shape == (1, 1500, shape.d_model) and o.dtype == torch.bfloat16 and o.is_cuda for o in outs)
    full = omel.log_mel(audio, shape.n_mels)
    content = full.shape[1] - 1
    want = [omel.pad_or_trim(full[:, k * 3000: min((k + 1) * 3000, content)]) for k in range(3)]
    for got_seg, want_seg in zip(model.seen_segments, want):
        assert got_seg.dtype == np.float32 and np.abs(got_seg - want_seg).max() <= 1e-4
    ref = oenc.encoder_forward(np.stack(want), weights, shape)
    cmp = oenc.compare(torch.cat(outs).cpu(), ref)
    assert cmp["cosine"] >= 0.999 and cmp["min_row_cosine"] >= 0.999 and cmp["max_abs"] <= 0.12, cmp
    # error behaviour the reference relies on: any Exception per chunk -> ChunkResult(success=False) (ref: :355-365)
    with pytest.raises(ValueError, match="Invalid input features shape"):
        model.encode(np.zeros((shape.n_mels + 1, 3000), np.float32))
