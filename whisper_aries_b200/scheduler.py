"""Chunk scheduler: shards 30-second windows over the GPUs of one box.

The reference cuts a file into work items, pushes them on a ``queue.Queue`` that N worker threads (one model replica
each) drain, gathers ``ChunkResult``s on a second queue and sorts them by ``chunk_id``
(ref: final_optimized_transcriber.py:27-47 ChunkWork/ChunkResult, :256-299 worker loop, :443-504 enqueue / collect /
sort); its older variant places replica ``i`` on GPU ``i % gpu_count`` ("Yasmeen's code/complete_fixed_whisper.py":
179-184).  Windows of the mel+encoder path all cost the same, so here the partition is static and contiguous
(``ceil(n / G)`` windows per GPU), every shard writes its rows straight into one preallocated output at
``offset = window index`` (order is restored by construction, no sort), and a failing shard is reported as
``ChunkResult(success=False, error=...)`` without taking the others down (ref: :355-365).  There is no collective:
the only cross-device step is this host-side gather.

Two launch styles share the partition function:
  * one process, one worker thread per GPU  -> ``ChunkScheduler`` (mirrors the reference's threads);
  * one process per GPU under torchrun      -> ``partition_windows(n, world_size)[rank]`` (bench.py --gpus N)."""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

N_SAMPLES = 480000


@dataclass
class ChunkWork:
    """One shard: windows [start, stop) of the call, owned by worker ``worker_id``."""
    chunk_id: int
    start: int
    stop: int
    worker_id: int = 0


@dataclass
class ChunkResult:
    chunk_id: int
    start: int
    stop: int
    worker_id: int
    success: bool = True
    error: Optional[str] = None
    processing_time: float = 0.0

    @property
    def n_windows(self) -> int:
        return self.stop - self.start


def partition_windows(n_windows: int, n_parts: int) -> list[tuple[int, int]]:
    """Contiguous blocks of ``ceil(n / parts)`` windows; trailing parts may be empty.  120 windows over 8 -> 15 each."""
    if n_parts <= 0:
        raise ValueError("n_parts must be positive")
    if n_windows < 0:
        raise ValueError("n_windows must be >= 0")
    per = -(-n_windows // n_parts) if n_windows else 0
    return [(min(i * per, n_windows), min((i + 1) * per, n_windows)) for i in range(n_parts)]


def split_into_windows(pcm: np.ndarray, n_samples: int = N_SAMPLES) -> np.ndarray:
    """1-D f32 PCM -> ``[ceil(len / n_samples), n_samples]``, the last window zero-padded (what ``pad_or_trim`` of the
    features amounts to for the final 30-s segment)."""
    x = np.asarray(pcm, dtype=np.float32).reshape(-1)
    n_win = max(1, -(-x.shape[0] // n_samples))
    if x.shape[0] == n_win * n_samples:
        return x.reshape(n_win, n_samples)
    out = np.zeros((n_win, n_samples), dtype=np.float32)
    out.reshape(-1)[: x.shape[0]] = x
    return out


def plan_reference_chunks(n_samples: int, chunk_length_s: float = 180.0, overlap_s: float = 5.0,
                          sample_rate: int = 16000) -> list[tuple[int, int]]:
    """The reference's work items as sample ranges of ONE PCM buffer.

    The reference cuts a file into ``ceil(duration / chunk_length)`` chunks of ``chunk_length + overlap`` seconds and
    COPIES each into its ``ChunkWork`` (ref: final_optimized_transcriber.py:422-449; ``get_chunk`` -> ``.astype`` copy at
    :126-135).  Here a chunk is just ``(start_sample, end_sample)`` into the caller's buffer; ``chunk_windows`` below
    turns it into 30-s window views without copying."""
    if n_samples <= 0:
        return []
    if chunk_length_s <= 0 or overlap_s < 0:
        raise ValueError("chunk_length_s must be > 0 and overlap_s >= 0")
    duration = n_samples / sample_rate
    total = int(np.ceil(duration / chunk_length_s))
    out = []
    for chunk_id in range(total):
        start_sec = chunk_id * chunk_length_s
        end_sec = min(start_sec + chunk_length_s + overlap_s, duration)
        a, b = max(0, int(start_sec * sample_rate)), min(n_samples, int(end_sec * sample_rate))
        out.append((a, b))
    return out


def chunk_windows(pcm: np.ndarray, start: int, stop: int, n_samples: int = N_SAMPLES) -> np.ndarray:
    """30-s windows ``[k, n_samples]`` of ``pcm[start:stop]``: a strided VIEW of the caller's buffer for the full
    windows (no copy, so a pinned buffer stays pinned); only a ragged last window is copied into a zero-padded row."""
    x = np.asarray(pcm).reshape(-1)[start:stop]
    full, rest = divmod(x.shape[0], n_samples)
    if rest == 0:
        return x.reshape(full, n_samples) if full else np.zeros((0, n_samples), x.dtype)
    tail = np.zeros((1, n_samples), dtype=x.dtype)
    tail[0, :rest] = x[full * n_samples:]
    if full == 0:
        return tail
    return np.concatenate([x[: full * n_samples].reshape(full, n_samples), tail])      # (copy: ragged input only)


# A worker consumes windows [start, stop) of `windows` and writes rows [start, stop) of `out`.
Worker = Callable[[np.ndarray, int, int, object], None]


@dataclass
class ChunkScheduler:
    """Runs one worker thread per device over a static block partition and gathers in window order."""
    workers: Sequence[Worker]
    results: list = field(default_factory=list)

    def run(self, windows, out) -> list[ChunkResult]:
        n = int(windows.shape[0])
        parts = partition_windows(n, len(self.workers))
        works = [ChunkWork(i, a, b, i) for i, (a, b) in enumerate(parts)]
        results: list[Optional[ChunkResult]] = [None] * len(works)

        def body(w: ChunkWork):
            t0 = time.perf_counter()
            res = ChunkResult(w.chunk_id, w.start, w.stop, w.worker_id)
            try:
                if w.stop > w.start:
                    self.workers[w.worker_id](windows, w.start, w.stop, out)
            except Exception as exc:                      # one bad shard must not sink the call (ref: :355-365)
                res.success = False
                res.error = f"{type(exc).__name__}: {exc}"
            res.processing_time = time.perf_counter() - t0
            results[w.chunk_id] = res

        threads = [threading.Thread(target=body, args=(w,), daemon=True, name=f"aries-shard-{w.chunk_id}")
                   for w in works]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        self.results = [r for r in results if r is not None]
        return self.results


def gpu_worker(model, micro_batch: int = 16) -> Worker:
    """Worker for one GPU: pinned staging, H2D on a side stream double-buffered against compute, fused
    PCM -> mel -> encoder on the device, D2H of the bf16 states into the caller's (pinned) output rows.
    ``windows`` may be float32 or int16 (ffmpeg's pcm_s16le, ref: utils.py:116): int16 halves the H2D bytes and is
    converted on the GPU."""
    import torch

    dev = model.encoder.device
    d, t = model.shape.d_model, model.shape.n_ctx
    state = {}

    def run(windows, start: int, stop: int, out) -> None:
        torch.cuda.set_device(dev)
        n_s = int(windows.shape[1])
        wt = windows if isinstance(windows, torch.Tensor) else torch.from_numpy(windows)
        if wt.dtype not in (torch.float32, torch.int16):
            raise ValueError("windows must be float32 or int16 PCM")
        if "bufs" not in state or state["n_s"] != n_s or state["dtype"] != wt.dtype:
            state["n_s"], state["dtype"] = n_s, wt.dtype
            state["bufs"] = [torch.empty((micro_batch, n_s), dtype=wt.dtype, device=dev) for _ in range(2)]
            state["outs"] = [torch.empty((micro_batch, t, d), dtype=torch.bfloat16, device=dev) for _ in range(2)]
            state["copy"] = torch.cuda.Stream(dev)
            state["comp"] = torch.cuda.Stream(dev)
            state["h2d_done"] = [torch.cuda.Event() for _ in range(2)]
            state["comp_done"] = [torch.cuda.Event() for _ in range(2)]
            state["d2h_done"] = [torch.cuda.Event() for _ in range(2)]
        copy, comp = state["copy"], state["comp"]
        ot = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
        if ot.dtype != torch.bfloat16:
            ot = ot.view(torch.bfloat16)
        steps = list(range(start, stop, micro_batch))
        for k, s in enumerate(steps):
            e = min(s + micro_batch, stop)
            i = k & 1
            with torch.cuda.stream(copy):
                if k >= 2:
                    copy.wait_event(state["comp_done"][i])      # buffer i is free once step k-2 consumed it
                state["bufs"][i][: e - s].copy_(wt[s:e], non_blocking=True)
                state["h2d_done"][i].record(copy)
            with torch.cuda.stream(comp):
                comp.wait_event(state["h2d_done"][i])
                if k >= 2:
                    comp.wait_event(state["d2h_done"][i])       # out buffer i drained by step k-2's D2H
                model.encode_audio(state["bufs"][i][: e - s], out=state["outs"][i][: e - s])
                state["comp_done"][i].record(comp)
            with torch.cuda.stream(copy):
                copy.wait_event(state["comp_done"][i])
                ot[s:e].copy_(state["outs"][i][: e - s], non_blocking=True)
                state["d2h_done"][i].record(copy)
        copy.synchronize()
        comp.synchronize()

    return run


def gpu_transcribe_worker(model, prompt: Sequence[int], max_length: int = 448, micro_batch: int = 64, **generate_kw) -> Worker:
    """Worker for one GPU that goes all the way to token ids: PCM -> log-mel -> encoder -> greedy
    ``generate`` with the encoder states staying in HBM -- the 3.84 MB per window that ``gpu_worker`` gathers to the host
    (the first scaling risk SURVEY.md 8e names) becomes ``max_length`` int32 per window.

    ``micro_batch`` windows are encoded and decoded together: a decode step reads the decoder's 1.6 GB of weights once
    whatever the batch, so 64 windows (4.5 ms per token step) cost 3.4x less per window than 8 (2.1 ms).

    ``out`` is an int32 array / tensor ``[n_windows, max_length + 1]``: column 0 = number of sampled ids (before EOT),
    columns 1.. = the prompt followed by the sampled ids and EOT padding (what ``aries_decoder_generate`` returns)."""
    import torch

    if model.decoder is None:
        raise ValueError("gpu_transcribe_worker needs a WhisperModel built with decoder weights (decoder_shape=...)")
    dev = model.encoder.device
    prompt = [int(t) for t in prompt]
    P = len(prompt)
    eot = model.decoder.tokens.eot
    state = {}

    def run(windows, start: int, stop: int, out) -> None:
        torch.cuda.set_device(dev)
        n_s = int(windows.shape[1])
        wt = windows if isinstance(windows, torch.Tensor) else torch.from_numpy(windows)
        if wt.dtype not in (torch.float32, torch.int16):
            raise ValueError("windows must be float32 or int16 PCM")
        if state.get("n_s") != n_s or state.get("dtype") != wt.dtype:
            state["n_s"], state["dtype"] = n_s, wt.dtype
            state["buf"] = torch.empty((micro_batch, n_s), dtype=wt.dtype, device=dev)     # int16 is scaled on the GPU
        ot = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
        L = min(int(max_length), model.decoder.shape.n_text_ctx)
        if ot.shape[1] != L + 1 or ot.dtype != torch.int32:
            raise ValueError(f"out must be int32 [n_windows, {L + 1}]")
        for s in range(start, stop, micro_batch):
            e = min(s + micro_batch, stop)
            buf = state["buf"][: e - s]
            buf.copy_(wt[s:e], non_blocking=True)
            enc = model.encode_audio(buf)
            res = model.decoder.generate(enc, [prompt] * (e - s), max_length=L, **generate_kw)
            for i, r in enumerate(res):
                ids = r.sequences_ids[0]
                row = ot[s + i]
                row[0] = len(ids)
                row[1:1 + P] = torch.tensor(prompt, dtype=torch.int32)
                row[1 + P:1 + P + len(ids)] = torch.tensor(ids, dtype=torch.int32) if ids else torch.empty(0, dtype=torch.int32)
                row[1 + P + len(ids):] = eot

    return run
