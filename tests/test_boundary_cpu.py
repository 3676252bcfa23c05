"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; host logic of the shim and scheduler.
No compute calls here (no GPU in this container)."""
import ctypes
import dataclasses
import os
import re

import time

import numpy as np
import pytest

from oracle import logmel as omel, synth as osynth
from whisper_aries_b200 import FeatureExtractor, _lib, partition_windows, synthetic
from whisper_aries_b200.scheduler import ChunkScheduler, split_into_windows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("aries_b200.h", "aries_b200_test.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        names |= set(re.findall(r"ARIES_API\s+[\w\s\*]+?\b(aries_\w+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 22
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    assert names == set(_lib.PROTOTYPES), "ctypes prototypes and the headers disagree"
    assert lib.aries_abi_version() == 101


def test_no_cpu_fallback_when_no_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.aries_init(0, ctypes.byref(h))
    assert rc == _lib.ARIES_ECUDA and "no CPU fallback" in _lib.last_error()
    fe = FeatureExtractor(feature_size=80)          # host-side construction works ...
    with pytest.raises(RuntimeError):               # ... but computing without a GPU fails loudly
        fe(np.zeros(16000, np.float32))
    with pytest.raises(ValueError):
        FeatureExtractor(device="cpu")


def test_pure_host_entry_points():
    lib = _lib.load()
    assert lib.aries_logmel_num_frames(480000, 160) == 3001
    assert lib.aries_logmel_num_frames(1234, 160) == 8
    assert lib.aries_logmel_num_frames(8000, 0) == 50
    assert lib.aries_encoder_workspace_bytes(None, 4) == 0
    assert lib.aries_logmel_destroy(None) == 0 and lib.aries_encoder_destroy(None) == 0 and lib.aries_destroy(None) == 0
    # row f1 / f3 entry points: NULL handles are rejected with the state error, never dereferenced
    assert lib.aries_decoder_destroy(None) == 0
    out = (ctypes.c_float * 5)()
    assert lib.aries_decoder_last_stats(None, out, 5) == _lib.ARIES_ESTATE and "decoder handle" in _lib.last_error()
    assert lib.aries_decoder_generate(None, None, 1, None, 1, None, None, None, None, None, None) == _lib.ARIES_ESTATE
    assert lib.aries_pcm_s16_to_f32(None, None, None, 0, None) == _lib.ARIES_ESTATE


def test_decoder_shapes_and_token_ids():
    """Row f1 host constants: Whisper's special-token layout for the two multilingual vocabularies and the decoder dims."""
    t = synthetic.WhisperTokens.for_vocab(51866)
    assert (t.eot, t.sot, t.transcribe, t.no_speech, t.no_timestamps, t.timestamp_begin) == (50257, 50258, 50360, 50363, 50364, 50365)
    assert 51866 - t.timestamp_begin == 1501                       # <|0.00|> .. <|30.00|> in 0.02-s steps
    t = synthetic.WhisperTokens.for_vocab(51865)
    assert (t.eot, t.sot, t.transcribe, t.no_speech, t.no_timestamps, t.timestamp_begin) == (50257, 50258, 50359, 50362, 50363, 50364)
    s = synthetic.DEC_SHAPES["large-v3"]
    assert (s.vocab, s.d_model, s.n_heads, s.n_layers, s.d_ffn, s.n_text_ctx, s.n_audio_ctx) == (51866, 1280, 20, 32, 5120, 448, 1500)
    w = synthetic.decoder_weights(synthetic.DEC_SHAPES["micro"], 1, tied=True)
    assert "decoder/projection/weight" not in w and w["decoder/layer_0/attention/linear_1/weight"].shape == (256, 128)
    assert not w["decoder/layer_0/self_attention/linear_0/bias"][128:256].any()      # Whisper's key projection has no bias


def test_feature_extractor_surface_matches_upstream():
    fe = FeatureExtractor(feature_size=128)
    assert (fe.n_fft, fe.hop_length, fe.chunk_length, fe.n_samples, fe.nb_max_frames) == (400, 160, 30, 480000, 3000)
    assert fe.time_per_frame == 0.01 and fe.sampling_rate == 16000
    assert fe.mel_filters.shape == (128, 201) and fe.mel_filters.dtype == np.float32
    for n in (80, 128):
        assert np.array_equal(FeatureExtractor.get_mel_filters(16000, 400, n), omel.mel_filterbank(n))
    with pytest.raises(ValueError):
        FeatureExtractor(n_fft=512)


def test_synthetic_generators_match_the_oracles():
    assert np.array_equal(synthetic.window_signal(5, 4000), osynth.window_signal(5, 4000))
    a, b = synthetic.encoder_weights(synthetic.SHAPES["micro"]), osynth.encoder_weights(osynth.SHAPES["micro"])
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    for k in synthetic.SHAPES:
        assert synthetic.SHAPES[k].flops_per_window == osynth.SHAPES[k].flops_per_window


def test_partition_is_contiguous_and_complete():
    assert partition_windows(120, 8) == [(i * 15, i * 15 + 15) for i in range(8)]
    assert partition_windows(120, 4)[-1] == (90, 120)
    for n in (0, 1, 5, 63, 64, 121):
        for g in (1, 2, 3, 8):
            parts = partition_windows(n, g)
            assert len(parts) == g and parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts if b > a or n == 0) <= max(1, -(-n // g))
    with pytest.raises(ValueError):
        partition_windows(4, 0)


def test_split_into_windows_zero_pads_the_tail():
    x = np.arange(10, dtype=np.float32) + 1
    w = split_into_windows(x, 4)
    assert w.shape == (3, 4) and w[2].tolist() == [9, 10, 0, 0]
    assert split_into_windows(x[:8], 4).shape == (2, 4)


def test_scheduler_with_fake_devices_keeps_order_and_isolates_failures():
    windows = np.arange(10 * 3, dtype=np.float32).reshape(10, 3)
    out = np.zeros((10, 2), np.float32)
    seen = []

    def make(worker_id, fail=False):
        def run(w, start, stop, o):
            seen.append((worker_id, start, stop))
            if fail:
                raise RuntimeError("device lost")
            o[start:stop, 0] = w[start:stop].sum(axis=1)
            o[start:stop, 1] = worker_id
        return run

    res = ChunkScheduler([make(0), make(1, fail=True), make(2)]).run(windows, out)
    assert [r.chunk_id for r in res] == [0, 1, 2]
    assert [r.success for r in res] == [True, False, True] and "device lost" in res[1].error
    assert [(r.start, r.stop) for r in res] == [(0, 4), (4, 8), (8, 10)]
    assert np.array_equal(out[:4, 0], windows[:4].sum(axis=1)) and (out[4:8] == 0).all()
    assert out[8:, 1].tolist() == [2, 2]
    # more devices than windows: trailing shards are empty and succeed
    res = ChunkScheduler([make(i) for i in range(4)]).run(windows[:2], np.zeros((2, 2), np.float32))
    assert [r.n_windows for r in res] == [1, 1, 0, 0] and all(r.success for r in res)


def test_submit_pipelines_two_jobs_through_two_phase_workers():
    """``submit`` returns at once and ``Job.result()`` gathers; a worker exposing ``enqueue(...) -> finish()`` (what
    ``gpu_worker`` does with CUDA streams) gets the SECOND job enqueued before the first one is finished, finishes them
    in order, and a failure in either phase lands in that job's ChunkResult only."""
    import threading
    from whisper_aries_b200 import Job
    events, gate = [], threading.Event()

    def make(worker_id):
        def enqueue(w, start, stop, o, lengths=None):
            events.append(("enqueue", int(w[0, 0]), worker_id))
            if w[0, 0] == 300:
                raise RuntimeError("bad input")

            def finish():
                if w[0, 0] == 100:
                    gate.wait(5.0)                       # job A stays in flight until job B has been enqueued
                if w[0, 0] == 400:
                    raise RuntimeError("device lost in flight")
                o[start:stop] = w[start:stop] + 1
                events.append(("finish", int(w[0, 0]), worker_id))
            return finish

        def run(w, start, stop, o, lengths=None):
            enqueue(w, start, stop, o, lengths=lengths)()
        run.enqueue = enqueue
        return run

    sched = ChunkScheduler([make(0)])
    a, b = np.full((2, 1), 100.0, np.float32), np.full((2, 1), 200.0, np.float32)
    oa, ob = np.zeros_like(a), np.zeros_like(b)
    ja = sched.submit(a, oa)
    jb = sched.submit(b, ob)
    assert isinstance(ja, Job)
    for _ in range(200):                                 # the worker thread enqueues B while A is still unfinished
        if ("enqueue", 200, 0) in events:
            break
        time.sleep(0.01)
    assert ("enqueue", 200, 0) in events and ("finish", 100, 0) not in events
    gate.set()
    assert all(r.success for r in ja.result()) and all(r.success for r in jb.result())
    assert (oa == 101).all() and (ob == 201).all()
    assert events.index(("finish", 100, 0)) < events.index(("finish", 200, 0))
    assert ja.result() is ja.result()                    # gathered once
    # failures: at enqueue time and in flight
    c, d_ = np.full((1, 1), 300.0, np.float32), np.full((1, 1), 400.0, np.float32)
    rc = sched.submit(c, np.zeros_like(c)).result()
    rd = sched.submit(d_, np.zeros_like(d_)).result()
    assert not rc[0].success and "bad input" in rc[0].error
    assert not rd[0].success and "in flight" in rd[0].error
    assert all(r.success for r in sched.run(b, ob))      # the worker thread survives both
    sched.close()


def test_reference_chunk_plan_and_zero_copy_windows():
    """Row f3: the reference's 180 s + 5 s work items (ref: final_optimized_transcriber.py:422-449) as sample ranges,
    and 30-s windows as views of the caller's buffer."""
    from whisper_aries_b200 import chunk_windows, plan_reference_chunks
    sr = 16000
    n = int(400.5 * sr)                                            # 400.5 s -> ceil(400.5 / 180) = 3 chunks
    plan = plan_reference_chunks(n)
    assert plan == [(0, 185 * sr), (180 * sr, 365 * sr), (360 * sr, n)]
    assert plan_reference_chunks(0) == [] and plan_reference_chunks(10 * sr) == [(0, 10 * sr)]
    with pytest.raises(ValueError):
        plan_reference_chunks(n, chunk_length_s=0)
    pcm = np.arange(n, dtype=np.float32)
    w = chunk_windows(pcm, 0, 6 * 480000)
    assert w.shape == (6, 480000) and np.shares_memory(w, pcm) and w[5, -1] == pcm[6 * 480000 - 1]
    a, b = plan[0]                                                  # 185 s = 6 full windows + 5 s
    w = chunk_windows(pcm, a, b)
    assert w.shape == (7, 480000) and w[6, 5 * sr - 1] == pcm[b - 1] and not w[6, 5 * sr:].any()
    assert np.array_equal(w[:6].reshape(-1), pcm[: 6 * 480000])
    assert chunk_windows(pcm, 5, 5).shape == (0, 480000)
    assert chunk_windows(pcm, 0, 100).shape == (1, 480000)


def test_dynamic_queue_self_schedules_and_threads_persist():
    """The reference's self-scheduling (ref: final_optimized_transcriber.py:243-246, :256-299): items on ONE shared queue,
    a slow worker takes fewer of them, results come back in window order, and the worker threads survive ``run``."""
    import threading
    import time
    windows = np.arange(24 * 2, dtype=np.float32).reshape(24, 2)
    taken = {0: 0, 1: 0}
    idents = {0: set(), 1: set()}

    def make(worker_id, delay):
        def run(w, start, stop, o):
            time.sleep(delay)
            taken[worker_id] += stop - start
            idents[worker_id].add(threading.get_ident())
            o[start:stop, 0] = w[start:stop].sum(axis=1)
        return run

    sched = ChunkScheduler([make(0, 0.001), make(1, 0.05)], policy="dynamic", chunk=2)
    for _ in range(2):
        out = np.zeros((24, 1), np.float32)
        res = sched.run(windows, out)
        assert [r.chunk_id for r in res] == list(range(12)) and all(r.success for r in res)
        assert [(r.start, r.stop) for r in res] == [(i, i + 2) for i in range(0, 24, 2)]
        assert np.array_equal(out[:, 0], windows.sum(axis=1))
    assert taken[0] > taken[1] > 0 and taken[0] + taken[1] == 48
    assert len(idents[0]) == 1 and len(idents[1]) == 1           # the same two threads served both calls
    sched.close()
    with pytest.raises(RuntimeError):
        sched.run(windows, np.zeros((24, 1), np.float32))
    assert [(w.start, w.stop) for w in ChunkScheduler([make(0, 0)] * 2, policy="dynamic").plan(17)] == \
        [(0, 3), (3, 6), (6, 9), (9, 12), (12, 15), (15, 17)]
    with pytest.raises(ValueError):
        ChunkScheduler([], policy="static")
    with pytest.raises(ValueError):
        ChunkScheduler([make(0, 0)], policy="round-robin")


def test_collector_timeout_keeps_waiting_while_a_worker_lives_and_a_failed_item_does_not_stop_its_worker():
    import time
    windows = np.zeros((4, 1), np.float32)

    def slow(w, start, stop, o):
        if start == 0:
            time.sleep(0.3)                                       # three collector timeouts pass; the thread is alive
        if start == 1:
            raise ValueError("bad window")
        o[start:stop] = 1

    with ChunkScheduler([slow], policy="dynamic", chunk=1, result_timeout=0.1) as sched:
        out = np.zeros((4, 1), np.float32)
        res = sched.run(windows, out)
    assert [r.success for r in res] == [True, False, True, True] and "bad window" in res[1].error
    assert out[:, 0].tolist() == [1, 0, 1, 1] and res[0].processing_time >= 0.3


def test_lengths_reach_the_worker_and_ragged_tail_keeps_its_length():
    from whisper_aries_b200 import chunk_windows
    from whisper_aries_b200.scheduler import _length_runs
    pcm = np.arange(10, dtype=np.float32) + 1
    w, n = chunk_windows(pcm, 0, 10, 4, return_lengths=True)
    assert w.shape == (3, 4) and n.tolist() == [4, 4, 2] and w[2].tolist() == [9, 10, 0, 0]
    w, n = chunk_windows(pcm, 0, 8, 4, return_lengths=True)
    assert n.tolist() == [4, 4]
    assert _length_runs(n, 0, 2, 4) == [(0, 2, 4)] and _length_runs(None, 1, 3, 4) == [(1, 3, 4)]
    assert _length_runs(np.array([4, 4, 2, 4]), 0, 4, 4) == [(0, 2, 4), (2, 3, 2), (3, 4, 4)]
    with pytest.raises(ValueError):
        _length_runs(np.array([4, 0]), 0, 2, 4)
    seen = []

    def worker(win, start, stop, o, lengths=None):
        seen.append(None if lengths is None else [int(v) for v in lengths[start:stop]])

    w, n = chunk_windows(pcm, 0, 10, 4, return_lengths=True)
    sched = ChunkScheduler([worker])
    assert all(r.success for r in sched.run(w, np.zeros((3, 1)), lengths=n)) and seen == [[4, 4, 2]]
    with pytest.raises(ValueError):
        sched.run(w, np.zeros((3, 1)), lengths=n[:2])
    sched.close()


def test_product_and_oracle_generators_are_the_same_bytes():
    """``whisper_aries_b200/synthetic.py`` (what bench.py / smoke() feed the kernels) and ``oracle/synth.py`` (what the
    oracle and the golden fixtures are built from) are kept as two files on purpose -- the product package must not
    import oracle/ -- so this test is what keeps them equal."""
    for seed in (0, 1, 2, 63):
        assert np.array_equal(synthetic.window_signal(seed, 48000), osynth.window_signal(seed, 48000))
    assert np.array_equal(synthetic.batch_signals(3, 5, 16000), osynth.batch_signals(3, 5, 16000))
    for name in ("micro", "tiny"):
        a, b = synthetic.encoder_weights(synthetic.SHAPES[name], 1234), osynth.encoder_weights(osynth.SHAPES[name], 1234)
        assert list(a) == list(b) and all(np.array_equal(a[k], b[k]) for k in a)
    da, db = synthetic.decoder_weights(synthetic.DEC_SHAPES["micro"], 32), osynth.decoder_weights(osynth.DEC_SHAPES["micro"], 32)
    assert list(da) == list(db) and all(np.array_equal(da[k], db[k]) for k in da)
    ta, tb = synthetic.WhisperTokens.for_vocab(51866), osynth.WhisperTokens.for_vocab(51866)
    assert dataclasses.asdict(ta) == dataclasses.asdict(tb)
    assert np.array_equal(synthetic.sinusoids(1500, 384), osynth.sinusoids(1500, 384))
