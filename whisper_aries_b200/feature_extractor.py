"""Drop-in for ``faster_whisper.feature_extractor.FeatureExtractor`` (v1.1.1) running on a B200.

Same constructor, attributes and ``__call__(waveform, padding=160, chunk_length=None)`` as upstream (SURVEY.md rows
a-1..a-4, 8b); the reference reaches it through ``model.transcribe`` (ref: final_optimized_transcriber.py:326).
The STFT / mel / log / clamp arithmetic runs in the hand-written sm_100a kernel behind ``aries_logmel_run``; only the
one-time filter-bank construction (f64 -> f32, as upstream does it) happens on the host.

Additive extensions: a ``[batch, n_samples]`` input returns ``[batch, n_mels, frames]``; a CUDA ``torch.Tensor``
input is consumed in place and a CUDA tensor is returned (no host round trip)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class FeatureExtractor:
    def __init__(self, feature_size: int = 80, sampling_rate: int = 16000, hop_length: int = 160,
                 chunk_length: int = 30, n_fft: int = 400, device="cuda:0"):
        if n_fft != 400 or hop_length != 160:
            raise ValueError("the B200 log-mel kernel is specialised for n_fft=400, hop_length=160 "
                             f"(every Whisper model); got n_fft={n_fft}, hop_length={hop_length}")
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.time_per_frame = hop_length / sampling_rate
        self.sampling_rate = sampling_rate
        self.mel_filters = self.get_mel_filters(sampling_rate, n_fft, n_mels=feature_size).astype("float32")
        self.device_index = _lib.device_index_of(device)
        self._handle = None

    # ------------------------------------------------------------------ host-side constants
    @staticmethod
    def get_mel_filters(sr, n_fft, n_mels=128):
        """Slaney-scale, Slaney-normalised triangular filters, float64 -> float32, shape [n_mels, n_fft // 2 + 1]."""
        n_mels = int(n_mels)
        fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
        min_mel, max_mel = 0.0, 45.245640471924965
        mels = np.linspace(min_mel, max_mel, n_mels + 2)
        f_sp = 200.0 / 3
        freqs = f_sp * mels
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
        fdiff = np.diff(freqs)
        ramps = freqs.reshape(-1, 1) - fftfreqs.reshape(1, -1)
        lower = -ramps[:-2] / np.expand_dims(fdiff[:-1], axis=1)
        upper = ramps[2:] / np.expand_dims(fdiff[1:], axis=1)
        weights = np.maximum(np.zeros_like(lower), np.minimum(lower, upper))
        enorm = 2.0 / (freqs[2:n_mels + 2] - freqs[:n_mels])
        weights *= np.expand_dims(enorm, axis=1)
        return weights.astype(np.float32)

    @property
    def feature_size(self) -> int:
        return int(self.mel_filters.shape[0])

    # ------------------------------------------------------------------ device handle
    def _mel(self):
        if self._handle is None:
            ctx = _lib.Context.get(self.device_index)
            filt = np.ascontiguousarray(self.mel_filters, dtype=np.float32)
            h = ctypes.c_void_p()
            _lib.check(ctx.lib.aries_logmel_create(ctx.handle, filt.shape[0], filt.ctypes.data, ctypes.byref(h)))
            self._ctx, self._handle = ctx, h
        return self._handle

    def close(self):
        if self._handle is not None:
            self._ctx.lib.aries_logmel_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launches(self) -> int:
        return 0 if self._handle is None else self._ctx.lib.aries_logmel_last_launches(self._handle)

    # ------------------------------------------------------------------ the call surface
    def __call__(self, waveform, padding: int = 160, chunk_length=None, *, frames_out=None):
        """Log-mel spectrogram: float32 ``[n_mels, (N + padding) // 160]`` for a 1-D input of N samples.

        ``frames_out`` (extension) stores that many frames per signal instead — 3000 gives exactly the
        ``pad_or_trim``-ed window ``encode`` consumes, with the clamp maximum still taken over every frame."""
        if chunk_length is not None:
            self.n_samples = chunk_length * self.sampling_rate
            self.nb_max_frames = self.n_samples // self.hop_length
        if padding < 0:
            raise ValueError("padding must be >= 0")
        try:
            import torch
            is_tensor = isinstance(waveform, torch.Tensor)
        except ImportError:                                      # pragma: no cover - torch is part of the image
            is_tensor = False
        if is_tensor and waveform.is_cuda:
            return self._call_device(waveform, padding, frames_out)
        if is_tensor:
            waveform = waveform.detach().cpu().numpy()
        x = np.asarray(waveform)
        if x.dtype != np.float32:
            x = x.astype(np.float32)
        squeeze = x.ndim == 1
        if squeeze:
            x = x[None]
        if x.ndim != 2 or x.shape[1] == 0:
            raise ValueError(f"waveform must be [n_samples] or [batch, n_samples] with n_samples > 0, got {x.shape}")
        x = np.ascontiguousarray(x)
        batch, n = x.shape
        mel = self._mel()
        lib = self._ctx.lib
        frames = int(lib.aries_logmel_num_frames(n, padding)) if frames_out is None else int(frames_out)
        out = np.empty((batch, self.feature_size, frames), dtype=np.float32)
        if frames > 0:
            _lib.check(lib.aries_logmel_run_host(mel, x.ctypes.data, batch, n, padding, out.ctypes.data, frames))
        return out[0] if squeeze else out

    def _call_device(self, waveform, padding, frames_out):
        import torch
        x = waveform
        if x.dtype != torch.float32:
            x = x.float()
        squeeze = x.dim() == 1
        if squeeze:
            x = x[None]
        if x.dim() != 2 or x.shape[1] == 0:
            raise ValueError(f"waveform must be [n_samples] or [batch, n_samples], got {tuple(x.shape)}")
        if x.stride(1) != 1:
            x = x.contiguous()
        if x.device.index != self.device_index:
            raise ValueError(f"waveform is on {x.device}, this extractor on cuda:{self.device_index}")
        batch, n = x.shape
        mel = self._mel()
        lib = self._ctx.lib
        frames = int(lib.aries_logmel_num_frames(n, padding)) if frames_out is None else int(frames_out)
        out = torch.empty((batch, self.feature_size, frames), dtype=torch.float32, device=x.device)
        if frames > 0:
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.aries_logmel_run(mel, x.data_ptr(), batch, n, x.stride(0), padding, out.data_ptr(), frames,
                                            stream))
        return out[0] if squeeze else out
