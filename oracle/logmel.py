"""numpy restatement of faster-whisper 1.1.1 ``FeatureExtractor`` (oracle; test infrastructure).

Follows SURVEY.md section 8 rows a-1 .. a-4:
  a-1 ``get_mel_filters``   Slaney-scale, Slaney-normalised triangular bank, f64 -> f32
  a-2 ``stft``              centre/reflect pad 200, 400-sample frames every 160, periodic Hann (f32),
                            rfft computed in f64 (numpy 1.26.4 behaviour) then rounded to complex64
  a-3 ``__call__``          pad 160 zeros -> stft -> drop last frame -> |.|^2 -> mel -> log10(clip 1e-10)
                            -> max(x, global_max - 8) -> (x + 4) / 4
  a-4 ``pad_or_trim``       window of exactly 3000 frames

The reference reaches this code only through ``model.transcribe(chunk_audio, ...)``
(ref: final_optimized_transcriber.py:326; warm-up at :189), which runs
``self.feature_extractor(audio)`` on the float32 PCM it is handed (ref: :306).

Everything here is written from the published algorithm; no upstream source is vendored.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_BINS = N_FFT // 2 + 1          # 201
CHUNK_SECONDS = 30
N_SAMPLES = CHUNK_SECONDS * SAMPLE_RATE   # 480000
NB_MAX_FRAMES = N_SAMPLES // HOP          # 3000
# Slaney mel of 8000 Hz: 15 + 27*ln(8)/ln(6.4); upstream hard-codes the literal.
MAX_MEL_8K = 45.245640471924965


def mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """Row a-1. Returns f32 [n_mels, n_fft//2+1]."""
    bin_hz = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)                 # f64 [201]
    mel_pts = np.linspace(0.0, MAX_MEL_8K, int(n_mels) + 2)        # f64
    hz_per_mel = 200.0 / 3.0
    edge_hz = hz_per_mel * mel_pts                                  # linear region
    knee_hz = 1000.0
    knee_mel = knee_hz / hz_per_mel                                 # 15.0
    log_step = np.log(6.4) / 27.0
    in_log = mel_pts >= knee_mel
    edge_hz[in_log] = knee_hz * np.exp(log_step * (mel_pts[in_log] - knee_mel))

    widths = np.diff(edge_hz)                                       # [n_mels+1]
    dist = edge_hz[:, None] - bin_hz[None, :]                       # [n_mels+2, 201]
    rising = -dist[:-2] / widths[:-1, None]
    falling = dist[2:] / widths[1:, None]
    tri = np.maximum(0.0, np.minimum(rising, falling))
    tri *= (2.0 / (edge_hz[2:n_mels + 2] - edge_hz[:n_mels]))[:, None]   # Slaney area norm
    return tri.astype(np.float32)


def hann_periodic_f32(n_fft: int = N_FFT) -> np.ndarray:
    """``np.hanning(n_fft+1)[:-1]`` cast to f32 (row a-3)."""
    return np.hanning(n_fft + 1)[:-1].astype(np.float32)


def padded_signal(waveform: np.ndarray, padding: int = HOP, n_fft: int = N_FFT) -> np.ndarray:
    """Zero-pad ``padding`` samples at the END, *then* reflect-pad n_fft//2 both sides (rows a-2/a-3)."""
    x = np.asarray(waveform)
    if x.dtype != np.float32:
        x = x.astype(np.float32)
    if padding:
        x = np.pad(x, (0, padding))
    return np.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")


def stft_power(waveform: np.ndarray, padding: int = HOP) -> np.ndarray:
    """|STFT|^2 as f32 [201, (N+padding)//160] — last STFT frame already dropped."""
    y = padded_signal(waveform, padding)
    n_frames = 1 + (y.shape[-1] - N_FFT) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = y[idx] * hann_periodic_f32()[None, :]                 # f32 * f32 -> f32
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)         # f64 transform (numpy 1.26.4 upcasts)
    spec = spec.astype(np.complex64).T                              # [201, n_frames]
    mag = np.abs(spec[:, :-1])                                      # f32
    return mag ** 2


def log_mel(waveform: np.ndarray, n_mels: int = 80, padding: int = HOP,
            filters: np.ndarray | None = None) -> np.ndarray:
    """Rows a-1..a-3: f32 [n_mels, (N+padding)//160]."""
    if filters is None:
        filters = mel_filterbank(n_mels)
    power = stft_power(waveform, padding)
    mel = filters @ power
    x = np.log10(np.clip(mel, a_min=1e-10, a_max=None))
    x = np.maximum(x, x.max() - 8.0)
    return ((x + 4.0) / 4.0).astype(np.float32)


def pad_or_trim(features: np.ndarray, length: int = NB_MAX_FRAMES) -> np.ndarray:
    """Row a-4: last axis cut / zero-padded to ``length`` frames."""
    n = features.shape[-1]
    if n > length:
        return features[..., :length]
    if n < length:
        pad = [(0, 0)] * (features.ndim - 1) + [(0, length - n)]
        return np.pad(features, pad)
    return features


def log_mel_window(waveform: np.ndarray, n_mels: int) -> np.ndarray:
    """The exact array ``encode`` sees for one 30-s window: [n_mels, 3000]."""
    return np.ascontiguousarray(pad_or_trim(log_mel(waveform, n_mels)))
