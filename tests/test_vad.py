"""Row f4: the VAD front end (``vad_filter=True``, ref: final_optimized_transcriber.py:440).

CPU: the C-ABI state machine (``aries_vad_speech_timestamps``, host-only) against the oracle restatement of
faster-whisper 1.1.1 ``get_speech_timestamps`` -- segment boundaries must be IDENTICAL (integer sample indices) -- on
hand-made and on seeded random probability tracks, over the option space the reference can reach.
GPU: ``collect_chunks`` as a device gather (bit-exact vs numpy), the energy stand-in model, and ``encode_long(vad_filter=True)``
against oracle VAD -> oracle log-mel -> oracle encoder."""
import numpy as np
import pytest

from oracle import vad as ov
from whisper_aries_b200 import vad


def as_tuples(chunks):
    return [(int(c["start"]), int(c["end"])) for c in chunks]


def both(probs, n, **kw):
    got = vad.speech_timestamps_from_probs(probs, n, vad.VadOptions(**kw))
    want = ov.get_speech_timestamps(np.asarray(probs, np.float32), n, ov.VadOptions(**kw))
    assert as_tuples(got) == as_tuples(want), (kw, as_tuples(got)[:6], as_tuples(want)[:6])
    return as_tuples(got)


def test_known_answers_with_upstream_defaults():
    # 31.25 windows per second; defaults: threshold 0.5 / 0.35, min_silence 2000 ms (62.5 windows), pad 400 ms (6400 samples)
    n = 400 * 512
    p = np.zeros(400, np.float32)
    assert both(p, n) == []                                             # silence: nothing reaches the feature extractor
    p[:] = 0.9
    assert both(p, n) == [(0, n)]                                       # all speech: one chunk, clipped to the audio
    p[:] = 0.0
    p[100:200] = 0.9                                                    # one burst: padded by 6400 samples on both sides
    assert both(p, n) == [(100 * 512 - 6400, 200 * 512 + 6400)]
    p[230:300] = 0.9                                                    # 30 windows of silence < 62.5: stays ONE chunk
    assert both(p, n) == [(100 * 512 - 6400, 300 * 512 + 6400)]
    p[:] = 0.0
    p[10:60] = 0.9
    p[200:260] = 0.9                                                    # 140 windows of silence: two chunks, each padded
    assert both(p, n) == [(0, 60 * 512 + 6400), (200 * 512 - 6400, 260 * 512 + 6400)]
    # hysteresis: 0.4 is below the threshold but above neg_threshold 0.35 -> it does not end a chunk; the zeros after
    # window 300 do (100 windows > 62.5), at the first silent window
    p[:] = 0.0
    p[50:60] = 0.9
    p[60:300] = 0.4
    assert both(p, n) == [(50 * 512 - 6400, 300 * 512 + 6400)]
    # short silence between chunks (< 2 * pad): the gap is split in the middle
    q = np.zeros(400, np.float32)
    q[10:40] = 0.9
    q[50:80] = 0.9
    assert both(q, n, min_silence_duration_ms=100, speech_pad_ms=400) == [(0, 45 * 512), (45 * 512, 80 * 512 + 6400)]
    assert vad.VadOptions() == vad.VadOptions(0.5, None, 0, float("inf"), 2000, 400)
    assert int(vad._lib.load().aries_vad_num_windows(1024)) == 3 == ov.n_windows(1024)       # a FULL extra window
    assert int(vad._lib.load().aries_vad_num_windows(1000)) == 2 == ov.n_windows(1000)


@pytest.mark.parametrize("seed", range(12))
def test_random_tracks_match_the_oracle_exactly(seed):
    rng = np.random.default_rng(seed)
    n_win = int(rng.integers(1, 900))
    # piecewise-constant tracks with noise: long and short speech / silence runs, values around both thresholds
    probs = np.zeros(n_win, np.float32)
    i = 0
    while i < n_win:
        run = int(rng.integers(1, 120))
        level = rng.choice([0.02, 0.2, 0.34, 0.36, 0.45, 0.55, 0.8, 0.97])
        probs[i:i + run] = np.clip(level + 0.05 * rng.standard_normal(min(run, n_win - i)), 0, 1)
        i += run
    n = n_win * 512 - int(rng.integers(1, 512))
    for kw in ({}, {"min_silence_duration_ms": 500}, {"min_speech_duration_ms": 250, "speech_pad_ms": 30},
               {"max_speech_duration_s": 5.0, "min_silence_duration_ms": 300},
               {"max_speech_duration_s": 2.0, "min_silence_duration_ms": 100, "speech_pad_ms": 0},
               {"threshold": 0.3, "neg_threshold": 0.1}, {"threshold": 0.7, "min_silence_duration_ms": 0}):
        chunks = both(probs, n, **kw)
        assert all(0 <= a <= b <= n for a, b in chunks)
        assert all(chunks[k][1] <= chunks[k + 1][0] for k in range(len(chunks) - 1))    # ordered, non-overlapping


def test_speech_timestamps_map_restores_original_times():
    chunks = [{"start": 16000, "end": 48000}, {"start": 80000, "end": 96000}]
    m = vad.SpeechTimestampsMap(chunks, 16000)
    for t in (0.0, 0.5, 1.99, 2.0, 2.5, 2.99, 10.0):
        assert m.get_original_time(t) == ov.restore_time(chunks, t)
    assert m.get_original_time(0.0) == 1.0 and m.get_original_time(2.5) == 5.5


def test_bad_arguments_are_value_errors():
    with pytest.raises(ValueError):
        vad.speech_timestamps_from_probs(np.zeros(4, np.float32), 2048, vad.VadOptions(threshold=0.0))
    with pytest.raises(ValueError):
        vad.EnergyVad(width_db=0)
    with pytest.raises(ValueError):
        vad.get_speech_timestamps(np.zeros(10, np.float32), sampling_rate=8000, speech_probs=np.zeros(1))


# ---------------------------------------------------------------------------------------------------- GPU
def speechy(seed, seconds, sr=16000):
    """Bursts of a loud tone between stretches of near-silence: what an energy VAD can cut."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    x = (1e-4 * rng.standard_normal(n)).astype(np.float32)
    t = 0
    while t < n:
        on = int(rng.integers(sr // 2, 4 * sr))
        off = int(rng.integers(sr // 4, 5 * sr))
        seg = slice(t, min(t + on, n))
        x[seg] += (0.2 * np.sin(2 * np.pi * 300.0 * np.arange(seg.stop - seg.start) / sr)).astype(np.float32)
        t += on + off
    return x


@pytest.mark.gpu
def test_collect_chunks_gather_is_bit_exact():
    import torch
    rng = np.random.default_rng(3)
    x = rng.standard_normal(1_000_003).astype(np.float32)
    d = torch.from_numpy(x).cuda()
    for chunks in ([], [{"start": 0, "end": 1}], [{"start": 5, "end": 5}, {"start": 7, "end": 1000}],
                   [{"start": int(a), "end": int(a + l)} for a, l in zip(range(0, 900000, 30011), rng.integers(0, 30000, 30))],
                   [{"start": 0, "end": x.shape[0]}]):
        got = vad.collect_chunks(d, chunks).cpu().numpy()
        assert np.array_equal(got, ov.collect_chunks(x, chunks)), chunks[:2]
    with pytest.raises(ValueError):
        vad.collect_chunks(d, [{"start": 10, "end": x.shape[0] + 1}])
    with pytest.raises(ValueError):
        vad.collect_chunks(d, [{"start": 10, "end": 5}])


@pytest.mark.gpu
def test_energy_standin_and_vad_filtered_encode_long_match_the_oracle_pipeline():
    import torch
    from oracle import encoder as oenc, logmel as omel, synth as osynth
    from whisper_aries_b200 import WhisperModel, synthetic
    x = speechy(11, 75.0)
    d = torch.from_numpy(x).cuda()
    probs = vad.EnergyVad()(d).cpu().numpy()
    want_probs = ov.energy_probs(x)
    assert probs.shape == want_probs.shape and np.abs(probs - want_probs).max() <= 2e-3
    opts = dict(min_silence_duration_ms=500)
    # the state machine is fed the SAME probabilities on both sides: boundaries must be identical
    chunks = vad.get_speech_timestamps(d, speech_probs=probs, **opts)
    assert as_tuples(chunks) == as_tuples(ov.get_speech_timestamps(probs, x.shape[0], ov.VadOptions(**opts)))
    assert 2 <= len(chunks) and sum(c["end"] - c["start"] for c in chunks) < 0.8 * x.shape[0]
    shape = synthetic.SHAPES["micro"]
    w = synthetic.encoder_weights(shape, 1234)
    model = WhisperModel("micro", w, device="cuda", device_index=0)
    out, used = model.encode_long(d, vad_filter=True, vad_parameters=opts, speech_probs=probs)
    assert as_tuples(used) == as_tuples(chunks)
    kept = ov.collect_chunks(x, chunks)
    full = omel.log_mel(kept, shape.n_mels)
    content = full.shape[1] - 1
    n_win = -(-content // 3000)
    feats = np.stack([omel.pad_or_trim(full[:, k * 3000: min((k + 1) * 3000, content)]) for k in range(n_win)])
    assert out.shape == (n_win, 1500, shape.d_model)
    cmp = oenc.compare(out.cpu(), oenc.encoder_forward(feats, w, shape))
    assert cmp["cosine"] >= 0.999 and cmp["min_row_cosine"] >= 0.999 and cmp["max_abs"] <= 0.12, cmp
    # pure silence: no chunk, nothing is encoded
    quiet = torch.zeros(48000, device="cuda")
    out0, used0 = model.encode_long(quiet, vad_filter=True)
    assert used0 == [] and out0.shape[0] == 0
