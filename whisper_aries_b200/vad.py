"""``vad_filter=True`` in front of the hot path.

The reference always transcribes with ``vad_filter=True`` (ref: final_optimized_transcriber.py:440, whitelisted at :318),
so faster-whisper 1.1.1's ``transcribe`` runs, before the feature extractor,

    speech_chunks = get_speech_timestamps(audio, vad_parameters)          # Silero VAD network + the state machine
    audio = np.concatenate(collect_chunks(audio, speech_chunks))          # only speech reaches FeatureExtractor
    ... SpeechTimestampsMap(speech_chunks, 16000) restores the original times afterwards

This module mirrors that surface (same names, options and defaults).  The state machine runs behind the C ABI
(``aries_vad_speech_timestamps``) and ``collect_chunks`` is one device gather (``aries_collect_chunks``), so the filtered
PCM goes from HBM to the log-mel kernel without visiting the host.

**The speech-probability model is pluggable and upstream's is NOT included**: Silero VAD's trained weights ship inside
the faster-whisper wheel (ONNX) and do not exist offline, so there is nothing to pin a re-implementation on.  Pass
``speech_probs=`` (e.g. from a Silero model you own) or ``model=`` (any callable: 1-D CUDA float32 PCM -> per-512-sample
window probabilities); the default, ``EnergyVad``, is a labelled stand-in (log-energy through a logistic) that makes
the path runnable end to end -- its decisions are not Silero's."""
from __future__ import annotations

import bisect
import ctypes
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _lib

WINDOW_SIZE_SAMPLES = 512


@dataclass
class VadOptions:
    """``faster_whisper.vad.VadOptions`` (1.1.1 defaults)."""
    threshold: float = 0.5
    neg_threshold: Optional[float] = None
    min_speech_duration_ms: int = 0
    max_speech_duration_s: float = float("inf")
    min_silence_duration_ms: int = 2000
    speech_pad_ms: int = 400

    def _c(self) -> "_lib.VadOpts":
        mx = float(self.max_speech_duration_s)
        return _lib.VadOpts(float(self.threshold), -1.0 if self.neg_threshold is None else float(self.neg_threshold),
                            int(self.min_speech_duration_ms), 0.0 if not np.isfinite(mx) else mx,
                            int(self.min_silence_duration_ms), int(self.speech_pad_ms))


class EnergyVad:
    """Stand-in probability model (NOT Silero): ``p = sigmoid((20 log10(rms of the 512-sample window) - center_db) /
    width_db)`` computed on the device (``aries_vad_energy_probs``)."""

    def __init__(self, center_db: float = -40.0, width_db: float = 4.0):
        if width_db <= 0:
            raise ValueError("width_db must be > 0")
        self.center_db, self.width_db = float(center_db), float(width_db)

    def __call__(self, audio):
        import torch
        if not (isinstance(audio, torch.Tensor) and audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 1):
            raise ValueError("EnergyVad expects a 1-D CUDA float32 tensor")
        audio = audio.contiguous()
        ctx = _lib.Context.get(audio.device.index or 0)
        n_win = int(ctx.lib.aries_vad_num_windows(audio.numel()))
        probs = torch.empty(n_win, dtype=torch.float32, device=audio.device)
        stream = torch.cuda.current_stream(audio.device).cuda_stream
        _lib.check(ctx.lib.aries_vad_energy_probs(ctx.handle, audio.data_ptr(), audio.numel(), self.center_db,
                                                  self.width_db, probs.data_ptr(), stream))
        return probs


def speech_timestamps_from_probs(speech_probs, audio_length_samples: int, vad_options: Optional[VadOptions] = None) -> list:
    """The state machine + padding pass of upstream's ``get_speech_timestamps`` over given probabilities ->
    ``[{"start": sample, "end": sample}, ...]``.  Host-only (``aries_vad_speech_timestamps`` needs no GPU)."""
    opts = vad_options or VadOptions()
    if not opts.threshold > 0:
        raise ValueError("threshold must be > 0")
    probs = np.ascontiguousarray(np.asarray(speech_probs, dtype=np.float32).reshape(-1))
    lib = _lib.load()
    c_opts = opts._c()
    cap = max(16, probs.shape[0] // 2 + 2)
    starts = np.empty(cap, dtype=np.int64)
    ends = np.empty(cap, dtype=np.int64)
    n = ctypes.c_int(0)
    _lib.check(lib.aries_vad_speech_timestamps(probs.ctypes.data, probs.shape[0], int(audio_length_samples),
                                               ctypes.byref(c_opts), starts.ctypes.data, ends.ctypes.data, cap,
                                               ctypes.byref(n)))
    return [{"start": int(starts[i]), "end": int(ends[i])} for i in range(n.value)]


def get_speech_timestamps(audio, vad_options: Optional[VadOptions] = None, sampling_rate: int = 16000, *,
                          speech_probs=None, model: Optional[Callable] = None, **kwargs) -> list:
    """``faster_whisper.vad.get_speech_timestamps``.  ``audio``: 1-D float32 PCM (CUDA tensor, host tensor or numpy; host
    input is uploaded).  ``speech_probs`` overrides the model; ``model`` defaults to the ``EnergyVad`` stand-in."""
    import torch
    if sampling_rate != 16000:
        raise ValueError("the hot path runs at 16 kHz (as upstream's VAD does)")
    if vad_options is None:
        vad_options = VadOptions(**kwargs)
    n = int(audio.shape[0])
    if speech_probs is None:
        if n == 0:
            return []
        x = audio if isinstance(audio, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
        if not x.is_cuda:
            x = x.cuda()
        speech_probs = (model or EnergyVad())(x.float())
    if isinstance(speech_probs, torch.Tensor):
        speech_probs = speech_probs.detach().cpu().numpy()
    return speech_timestamps_from_probs(speech_probs, n, vad_options)


def collect_chunks(audio, chunks: list):
    """``faster_whisper.vad.collect_chunks`` + the ``np.concatenate`` upstream applies to its result: the speech chunks
    of a 1-D CUDA float32 tensor, concatenated by one device gather (stream-ordered, no host round trip)."""
    import torch
    if not (isinstance(audio, torch.Tensor) and audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 1):
        raise ValueError("collect_chunks expects a 1-D CUDA float32 tensor")
    audio = audio.contiguous()
    if not chunks:
        return torch.empty(0, dtype=torch.float32, device=audio.device)
    starts = np.ascontiguousarray([int(c["start"]) for c in chunks], dtype=np.int64)
    ends = np.ascontiguousarray([int(c["end"]) for c in chunks], dtype=np.int64)
    total = int((ends - starts).sum())
    out = torch.empty(max(total, 0), dtype=torch.float32, device=audio.device)
    ctx = _lib.Context.get(audio.device.index or 0)
    n = ctypes.c_int64(0)
    stream = torch.cuda.current_stream(audio.device).cuda_stream
    _lib.check(ctx.lib.aries_collect_chunks(ctx.handle, audio.data_ptr(), audio.numel(), starts.ctypes.data,
                                            ends.ctypes.data, len(chunks), out.data_ptr() if total > 0 else None,
                                            out.numel(), ctypes.byref(n), stream))
    return out[: n.value]


class SpeechTimestampsMap:
    """``faster_whisper.vad.SpeechTimestampsMap``: times on the filtered (concatenated) axis -> the original axis."""

    def __init__(self, chunks: list, sampling_rate: int, time_precision: int = 2):
        self.sampling_rate = sampling_rate
        self.time_precision = time_precision
        self.chunk_end_sample = []
        self.total_silence_before = []
        previous_end = 0
        silent_samples = 0
        for chunk in chunks:
            silent_samples += chunk["start"] - previous_end
            previous_end = chunk["end"]
            self.chunk_end_sample.append(chunk["end"] - silent_samples)
            self.total_silence_before.append(silent_samples / sampling_rate)

    def get_original_time(self, time: float, chunk_index: Optional[int] = None) -> float:
        if chunk_index is None:
            chunk_index = self.get_chunk_index(time)
        return round(self.total_silence_before[chunk_index] + time, self.time_precision)

    def get_chunk_index(self, time: float) -> int:
        sample = int(time * self.sampling_rate)
        return min(bisect.bisect(self.chunk_end_sample, sample), len(self.chunk_end_sample) - 1)
