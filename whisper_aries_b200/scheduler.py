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

import collections
import queue
import threading
import time
from dataclasses import dataclass
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
    """1-D f32 PCM -> ``[ceil(len / n_samples), n_samples]``, the last window zero-padded IN THE SAMPLE DOMAIN.

    Note that this is not what upstream's ``pad_or_trim`` does to a short final segment: faster-whisper zero-fills the
    missing *feature frames* (0.0), whereas zero PCM gives log10(1e-10) clamped to ``(max - 8 + 4) / 4`` in every padded
    frame.  To reproduce upstream on a ragged tail keep its true length: ``chunk_windows(..., return_lengths=True)``
    and pass ``lengths`` to ``ChunkScheduler.run`` -- the kernel then writes 0.0 after the last real frame."""
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


def chunk_windows(pcm: np.ndarray, start: int, stop: int, n_samples: int = N_SAMPLES, return_lengths: bool = False):
    """30-s windows ``[k, n_samples]`` of ``pcm[start:stop]``: a strided VIEW of the caller's buffer for the full
    windows (no copy, so a pinned buffer stays pinned); only a ragged last window is copied into a zero-padded row.

    ``return_lengths=True`` also returns the valid samples per window (int64 ``[k]``).  Workers given those lengths
    run the ragged tail with its true sample count, so its features are ``pad_or_trim(log_mel(tail))`` -- zeros after
    the last real frame, as upstream (faster-whisper ``pad_or_trim``) -- instead of the log-mel of zero PCM."""
    x = np.asarray(pcm).reshape(-1)[start:stop]
    full, rest = divmod(x.shape[0], n_samples)
    if rest == 0:
        w = x.reshape(full, n_samples) if full else np.zeros((0, n_samples), x.dtype)
        return (w, np.full(full, n_samples, np.int64)) if return_lengths else w
    tail = np.zeros((1, n_samples), dtype=x.dtype)
    tail[0, :rest] = x[full * n_samples:]
    w = tail if full == 0 else np.concatenate([x[: full * n_samples].reshape(full, n_samples), tail])   # (copy: ragged input only)
    if return_lengths:
        return w, np.array([n_samples] * full + [rest], np.int64)
    return w


def _length_runs(lengths, s: int, e: int, n_s: int):
    """Maximal runs ``(a, b, n)`` of windows [s, e) sharing one valid length ``n`` (``lengths=None``: one full run)."""
    if lengths is None:
        return [(s, e, n_s)]
    runs, a = [], s
    for i in range(s + 1, e + 1):
        if i == e or int(lengths[i]) != int(lengths[a]):
            n = int(lengths[a])
            if n <= 0 or n > n_s:
                raise ValueError(f"window {a}: length {n} outside 1..{n_s}")
            runs.append((a, i, n))
            a = i
    return runs


# A worker consumes windows [start, stop) of `windows` and writes rows [start, stop) of `out`.  A worker that
# understands ragged windows also accepts ``lengths=`` (valid samples per window, see ``chunk_windows``).
Worker = Callable[..., None]

_STOP = object()          # poison pill (ref: final_optimized_transcriber.py:268 ``if work_item is None: break``)


class ChunkScheduler:
    """One persistent worker thread per device; work items are handed out over queues and gathered in window order.

    ``policy="static"`` (default, what the mel+encoder path wants: every window costs the same): one contiguous block
    of ``ceil(n / G)`` windows per worker.  ``policy="dynamic"``: the job is cut into items of ``chunk`` windows on ONE
    shared ``queue.Queue`` that the workers drain as they become free -- the reference's self-scheduling
    (ref: final_optimized_transcriber.py:243-246 queues, :256-299 worker loop), which is what shards that finish at
    ragged times need (``gpu_transcribe_worker``: windows stop decoding when they emit EOT).

    The collector waits ``result_timeout`` seconds per result (ref: :470, 120 s) and gives up only when no worker
    thread is alive any more (ref: :483-490); a failing item becomes ``ChunkResult(success=False, error=...)`` and the
    worker moves on to its next item (ref: :280-293, :355-365).  Threads are started on the first ``run`` and stay
    alive across calls until ``close()`` (the reference spawns and joins them per file, ref: :367-388, :395-408)."""

    def __init__(self, workers: Sequence[Worker], *, policy: str = "static", chunk: Optional[int] = None,
                 result_timeout: float = 120.0):
        if policy not in ("static", "dynamic"):
            raise ValueError("policy must be 'static' or 'dynamic'")
        if not workers:
            raise ValueError("ChunkScheduler needs at least one worker")
        if chunk is not None and chunk <= 0:
            raise ValueError("chunk must be positive")
        self.workers = list(workers)
        self.policy = policy
        self.chunk = chunk
        self.result_timeout = float(result_timeout)
        self.results: list[ChunkResult] = []
        self._threads: list[threading.Thread] = []
        self._inboxes: list[queue.Queue] = []
        self._lock = threading.Lock()
        self._closed = False

    # ------------------------------------------------------------------ worker threads
    def _loop(self, worker_id: int, inbox: "queue.Queue") -> None:
        worker = self.workers[worker_id]
        enqueue = getattr(worker, "enqueue", None)       # two-phase worker: enqueue(...) -> finish()  (see gpu_worker)
        pending: "collections.deque" = collections.deque()

        def complete_oldest() -> None:
            res, results, finish, t0 = pending.popleft()
            try:
                finish()
            except Exception as exc:
                res.success = False
                res.error = f"{type(exc).__name__}: {exc}"
            res.processing_time = time.perf_counter() - t0
            results.put(res)

        while True:
            # With work in flight only LOOK for the next item: if one is already queued (a pipelined ``submit``) its H2D
            # and kernels are enqueued behind the current item's before this thread blocks on the current item's D2H.
            if pending:
                try:
                    job = inbox.get_nowait()
                except queue.Empty:
                    job = None
            else:
                job = inbox.get()
            if job is _STOP:
                while pending:
                    complete_oldest()
                return
            if job is not None:
                work, windows, out, lengths, results = job
                res = ChunkResult(work.chunk_id, work.start, work.stop, worker_id)
                t0 = time.perf_counter()
                try:
                    if work.stop > work.start:
                        kw = {} if lengths is None else {"lengths": lengths}
                        if enqueue is not None:
                            pending.append((res, results, enqueue(windows, work.start, work.stop, out, **kw), t0))
                            res = None
                        else:
                            worker(windows, work.start, work.stop, out, **kw)
                except Exception as exc:                  # one bad shard must not sink the call (ref: :355-365)
                    res.success = False
                    res.error = f"{type(exc).__name__}: {exc}"
                if res is not None:
                    res.processing_time = time.perf_counter() - t0
                    results.put(res)
            if pending and (job is None or len(pending) > 1):
                complete_oldest()

    def _ensure_threads(self) -> None:
        if self._closed:
            raise RuntimeError("ChunkScheduler is closed")
        if self._threads:
            return
        shared = queue.Queue() if self.policy == "dynamic" else None
        for i in range(len(self.workers)):
            inbox = shared if shared is not None else queue.Queue()
            t = threading.Thread(target=self._loop, args=(i, inbox), daemon=True, name=f"aries-worker-{i}")
            self._inboxes.append(inbox)
            self._threads.append(t)
            t.start()

    def close(self) -> None:
        """Stop the worker threads (poison pill per thread, as the reference's ``stop_workers``, ref: :395-408)."""
        with self._lock:
            if self._closed:
                return
            self._closed = True
            for inbox in self._inboxes:
                inbox.put(_STOP)
            for t in self._threads:
                t.join(timeout=5.0)
            self._threads, self._inboxes = [], []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ one job
    def plan(self, n_windows: int) -> list[ChunkWork]:
        """The work items of a job of ``n_windows`` windows under this scheduler's policy."""
        if self.policy == "static":
            return [ChunkWork(i, a, b, i) for i, (a, b) in enumerate(partition_windows(n_windows, len(self.workers)))]
        step = self.chunk or max(1, -(-n_windows // (4 * len(self.workers))))
        return [ChunkWork(i, a, min(a + step, n_windows), -1) for i, a in enumerate(range(0, n_windows, step))]

    def submit(self, windows, out, lengths=None) -> "Job":
        """Enqueue a job and return at once; ``Job.result()`` gathers it.  Two jobs may be in flight per scheduler: with
        two-phase workers (``gpu_worker``) the second job's H2D and kernels queue up behind the first's on the device,
        so the first job's D2H overlaps the second's compute -- across calls, not only across micro-batches.  The caller
        keeps ``windows`` and ``out`` of a job alive and untouched until its ``result()`` has returned."""
        n = int(windows.shape[0])
        if lengths is not None and len(lengths) != n:
            raise ValueError("lengths must hold one entry per window")
        with self._lock:
            self._ensure_threads()
            works = self.plan(n)
            results: "queue.Queue" = queue.Queue()        # per job: a late result of an abandoned job cannot leak in
            for w in works:
                inbox = self._inboxes[w.worker_id if self.policy == "static" else 0]
                inbox.put((w, windows, out, lengths, results))
        return Job(self, works, results)

    def _gather(self, works, results) -> list[ChunkResult]:
        got: dict[int, ChunkResult] = {}
        while len(got) < len(works):
            try:
                r = results.get(timeout=self.result_timeout)
                got[r.chunk_id] = r
            except queue.Empty:
                if not any(t.is_alive() for t in self._threads):
                    break                                  # ref: :483-490 "All workers stopped, breaking..."
        for w in works:
            if w.chunk_id not in got:
                got[w.chunk_id] = ChunkResult(w.chunk_id, w.start, w.stop, w.worker_id, success=False,
                                              error="no result: every worker thread has stopped")
        self.results = [got[w.chunk_id] for w in works]
        return self.results

    def run(self, windows, out, lengths=None) -> list[ChunkResult]:
        """Process every window of ``windows`` into the rows of ``out`` (one preallocated, ideally pinned, gather
        buffer: order is restored by construction).  ``lengths`` = valid samples per window (``chunk_windows(...,
        return_lengths=True)``); ``None`` means every window is full.  Returns when every row is on the host."""
        return self.submit(windows, out, lengths).result()


class Job:
    """A submitted job (``ChunkScheduler.submit``): ``result()`` blocks until every work item has reported and returns
    the ``ChunkResult`` list in window order (the same collector and time-out as ``run``)."""

    def __init__(self, scheduler: ChunkScheduler, works, results):
        self._scheduler, self._works, self._results, self._done = scheduler, works, results, None

    def result(self) -> list[ChunkResult]:
        if self._done is None:
            self._done = self._scheduler._gather(self._works, self._results)
        return self._done


def gpu_worker(model, micro_batch: int = 16) -> Worker:
    """Worker for one GPU: uploads, kernels and downloads on three streams, double-buffered (the upload of micro-batch
    k+1 and the download of k-1 overlap the kernels of k, within a call and -- through ``enqueue`` -- across calls), fused
    PCM -> mel -> encoder on the device, D2H of the bf16 states into the caller's (pinned) output rows.
    ``windows`` may be float32 or int16 (ffmpeg's pcm_s16le, ref: utils.py:116): int16 halves the H2D bytes and is
    converted on the GPU."""
    import torch

    dev = model.encoder.device
    d, t = model.shape.d_model, model.shape.n_ctx
    state = {}

    def enqueue(windows, start: int, stop: int, out, lengths=None):
        """Enqueue H2D, kernels and D2H of windows [start, stop) and return ``finish()``, which blocks until the rows are
        on the host.  Micro-batches alternate between two device buffer pairs ACROSS calls, so a second call enqueued
        before the first is finished overlaps with it exactly as consecutive micro-batches of one call do."""
        torch.cuda.set_device(dev)
        n_s = int(windows.shape[1])
        wt = windows if isinstance(windows, torch.Tensor) else torch.from_numpy(windows)
        if wt.dtype not in (torch.float32, torch.int16):
            raise ValueError("windows must be float32 or int16 PCM")
        if "bufs" not in state or state["n_s"] != n_s or state["dtype"] != wt.dtype:
            if "comp" in state:                           # a new shape: nothing of the old one may still be in flight
                for name in ("h2d", "comp", "d2h"):
                    state[name].synchronize()
            state["n_s"], state["dtype"] = n_s, wt.dtype
            state["bufs"] = [torch.empty((micro_batch, n_s), dtype=wt.dtype, device=dev) for _ in range(2)]
            state["outs"] = [torch.empty((micro_batch, t, d), dtype=torch.bfloat16, device=dev) for _ in range(2)]
            # one stream per direction: on a shared copy stream the upload of micro-batch k+1 queues up BEHIND the
            # download of micro-batch k, which waits for k's kernels -- nothing would overlap (round 2 finding)
            state["h2d"] = torch.cuda.Stream(dev)
            state["comp"] = torch.cuda.Stream(dev)
            state["d2h"] = torch.cuda.Stream(dev)
            state["h2d_done"] = [torch.cuda.Event() for _ in range(2)]
            state["comp_done"] = [torch.cuda.Event() for _ in range(2)]
            state["d2h_done"] = [torch.cuda.Event() for _ in range(2)]
            state["k"] = 0                                # micro-batches enqueued so far, over all calls
        h2d, comp, d2h = state["h2d"], state["comp"], state["d2h"]
        ot = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
        if ot.dtype != torch.bfloat16:
            ot = ot.view(torch.bfloat16)
        for s in range(start, stop, micro_batch):
            e = min(s + micro_batch, stop)
            k = state["k"]
            state["k"] = k + 1
            i = k & 1
            with torch.cuda.stream(h2d):
                if k >= 2:
                    h2d.wait_event(state["comp_done"][i])       # buffer i is free once micro-batch k-2 consumed it
                state["bufs"][i][: e - s].copy_(wt[s:e], non_blocking=True)
                state["h2d_done"][i].record(h2d)
            with torch.cuda.stream(comp):
                comp.wait_event(state["h2d_done"][i])
                if k >= 2:
                    comp.wait_event(state["d2h_done"][i])       # out buffer i drained by micro-batch k-2's D2H
                for a, b, n in _length_runs(lengths, s, e, n_s):      # a ragged tail runs with its true length
                    model.encode_audio(state["bufs"][i][a - s: b - s, :n], out=state["outs"][i][a - s: b - s])
                state["comp_done"][i].record(comp)
            with torch.cuda.stream(d2h):
                d2h.wait_event(state["comp_done"][i])
                ot[s:e].copy_(state["outs"][i][: e - s], non_blocking=True)
                state["d2h_done"][i].record(d2h)
        done = torch.cuda.Event()
        done.record(d2h)                                  # downloads are FIFO on their stream: this covers the whole call

        def finish() -> None:
            done.synchronize()

        return finish

    def run(windows, start: int, stop: int, out, lengths=None) -> None:
        enqueue(windows, start, stop, out, lengths=lengths)()

    run.enqueue = enqueue
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
    state = {}

    def run(windows, start: int, stop: int, out, lengths=None) -> None:
        torch.cuda.set_device(dev)
        n_s = int(windows.shape[1])
        wt = windows if isinstance(windows, torch.Tensor) else torch.from_numpy(windows)
        if wt.dtype not in (torch.float32, torch.int16):
            raise ValueError("windows must be float32 or int16 PCM")
        if state.get("n_s") != n_s or state.get("dtype") != wt.dtype:
            state["n_s"], state["dtype"] = n_s, wt.dtype
            # int16 is scaled on the GPU; two staging buffers: the next micro-batch's H2D overlaps this one's decode
            state["bufs"] = [torch.empty((micro_batch, n_s), dtype=wt.dtype, device=dev) for _ in range(2)]
            state["enc"] = torch.empty((micro_batch, model.shape.n_ctx, model.shape.d_model), dtype=torch.bfloat16, device=dev)
            state["copy"] = torch.cuda.Stream(dev)
            state["comp"] = torch.cuda.Stream(dev)
            state["h2d_done"] = [torch.cuda.Event() for _ in range(2)]
            state["comp_done"] = [torch.cuda.Event() for _ in range(2)]
        ot = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
        L = min(int(max_length), model.decoder.shape.n_text_ctx)
        if ot.shape[1] != L + 1 or ot.dtype != torch.int32:
            raise ValueError(f"out must be int32 [n_windows, {L + 1}]")
        on = ot.numpy()                                        # host rows are filled with one vectorised store each
        copy, comp = state["copy"], state["comp"]
        steps = list(range(start, stop, micro_batch))

        def upload(k):
            s = steps[k]
            e = min(s + micro_batch, stop)
            i = k & 1
            with torch.cuda.stream(copy):
                if k >= 2:
                    copy.wait_event(state["comp_done"][i])
                state["bufs"][i][: e - s].copy_(wt[s:e], non_blocking=True)
                state["h2d_done"][i].record(copy)

        if steps:
            upload(0)
        for k, s in enumerate(steps):
            e = min(s + micro_batch, stop)
            i = k & 1
            if k + 1 < len(steps):
                upload(k + 1)
            with torch.cuda.stream(comp):
                comp.wait_event(state["h2d_done"][i])
                enc = state["enc"][: e - s]
                for a, b, n in _length_runs(lengths, s, e, n_s):
                    model.encode_audio(state["bufs"][i][a - s: b - s, :n], out=enc[a - s: b - s])
                state["comp_done"][i].record(comp)
                toks, lens = model.decoder.generate_ids(enc, prompt, max_length=L, **generate_kw)   # syncs on `comp`
            on[s:e, 0] = lens
            on[s:e, 1:] = toks

    return run
