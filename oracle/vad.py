"""CPU restatement of faster-whisper 1.1.1 ``vad.py`` post-processing (oracle; test infrastructure -- SURVEY.md row f4).

The reference always transcribes with ``vad_filter=True`` (ref: final_optimized_transcriber.py:440, whitelisted at :318),
so upstream's ``transcribe`` runs ``get_speech_timestamps(audio, VadOptions())`` -> ``collect_chunks`` BEFORE the feature
extractor: the VAD decides which samples reach the log-mel kernel.  Upstream's pipeline is
    speech_probs = SileroVADModel(padded_audio)            # ONNX network, 512-sample windows  (NOT restated here)
    speeches     = the state machine below over speech_probs
    audio        = np.concatenate([audio[c["start"]:c["end"]] for c in speeches])
This file restates the state machine and ``collect_chunks`` / ``SpeechTimestampsMap`` exactly as published for 1.1.1
[unverified offline: faster_whisper is not installable here -- parity unpinned against the wheel; the defaults
threshold 0.5, neg_threshold = max(threshold - 0.15, 0.01), min_speech 0 ms, max_speech inf, min_silence 2000 ms,
speech_pad 400 ms, window 512 samples, 98 ms "min silence at max speech" are the 1.1.x values].  The Silero network's
weights ship inside the wheel and are not available offline; the probabilities are an INPUT here.
Pure-Python loops on purpose: it is the checker for small cases, never the product path."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

WINDOW = 512


@dataclass
class VadOptions:
    threshold: float = 0.5
    neg_threshold: float | None = None
    min_speech_duration_ms: int = 0
    max_speech_duration_s: float = float("inf")
    min_silence_duration_ms: int = 2000
    speech_pad_ms: int = 400


def n_windows(audio_len: int) -> int:
    """Upstream pads with ``window - len % window`` samples (a FULL extra window when len is a multiple of 512)."""
    return (audio_len + (WINDOW - audio_len % WINDOW)) // WINDOW


def get_speech_timestamps(speech_probs, audio_length_samples: int, opts: VadOptions | None = None,
                          sampling_rate: int = 16000) -> list[dict]:
    opts = opts or VadOptions()
    threshold = opts.threshold
    neg_threshold = opts.neg_threshold
    window_size_samples = WINDOW
    min_speech_samples = sampling_rate * opts.min_speech_duration_ms / 1000
    speech_pad_samples = sampling_rate * opts.speech_pad_ms / 1000
    max_speech_samples = sampling_rate * opts.max_speech_duration_s - window_size_samples - 2 * speech_pad_samples
    min_silence_samples = sampling_rate * opts.min_silence_duration_ms / 1000
    min_silence_samples_at_max_speech = sampling_rate * 98 / 1000
    if neg_threshold is None:
        neg_threshold = max(threshold - 0.15, 0.01)

    triggered = False
    speeches: list[dict] = []
    current_speech: dict = {}
    temp_end = 0                 # potential segment end (tolerates some silence)
    prev_end = next_start = 0    # potential limits when the maximum segment size is reached

    for i, speech_prob in enumerate(speech_probs):
        # numpy 1.26.4 (the reference's pin) compares a float32 scalar with a Python float in float64
        speech_prob = float(speech_prob)
        if (speech_prob >= threshold) and temp_end:
            temp_end = 0
            if next_start < prev_end:
                next_start = window_size_samples * i

        if (speech_prob >= threshold) and not triggered:
            triggered = True
            current_speech["start"] = window_size_samples * i
            continue

        if triggered and (window_size_samples * i) - current_speech["start"] > max_speech_samples:
            if prev_end:
                current_speech["end"] = prev_end
                speeches.append(current_speech)
                current_speech = {}
                if next_start < prev_end:        # previously reached silence and is still not speech
                    triggered = False
                else:
                    current_speech["start"] = next_start
                prev_end = next_start = temp_end = 0
            else:
                current_speech["end"] = window_size_samples * i
                speeches.append(current_speech)
                current_speech = {}
                prev_end = next_start = temp_end = 0
                triggered = False
                continue

        if (speech_prob < neg_threshold) and triggered:
            if not temp_end:
                temp_end = window_size_samples * i
            if (window_size_samples * i) - temp_end > min_silence_samples_at_max_speech:   # avoid cutting in very short silence
                prev_end = temp_end
            if (window_size_samples * i) - temp_end < min_silence_samples:
                continue
            else:
                current_speech["end"] = temp_end
                if (current_speech["end"] - current_speech["start"]) > min_speech_samples:
                    speeches.append(current_speech)
                current_speech = {}
                prev_end = next_start = temp_end = 0
                triggered = False
                continue

    if current_speech and (audio_length_samples - current_speech["start"]) > min_speech_samples:
        current_speech["end"] = audio_length_samples
        speeches.append(current_speech)

    for i, speech in enumerate(speeches):
        if i == 0:
            speech["start"] = int(max(0, speech["start"] - speech_pad_samples))
        if i != len(speeches) - 1:
            silence_duration = speeches[i + 1]["start"] - speech["end"]
            if silence_duration < 2 * speech_pad_samples:
                speech["end"] += int(silence_duration // 2)
                speeches[i + 1]["start"] = int(max(0, speeches[i + 1]["start"] - silence_duration // 2))
            else:
                speech["end"] = int(min(audio_length_samples, speech["end"] + speech_pad_samples))
                speeches[i + 1]["start"] = int(max(0, speeches[i + 1]["start"] - speech_pad_samples))
        else:
            speech["end"] = int(min(audio_length_samples, speech["end"] + speech_pad_samples))
    return speeches


def collect_chunks(audio: np.ndarray, chunks: list[dict]) -> np.ndarray:
    """What ``transcribe`` feeds the feature extractor when ``vad_filter=True``: the speech chunks, concatenated."""
    if not chunks:
        return np.array([], dtype=np.float32)
    return np.concatenate([audio[c["start"]: c["end"]] for c in chunks])


def restore_time(chunks: list[dict], time: float, sampling_rate: int = 16000, time_precision: int = 2) -> float:
    """``SpeechTimestampsMap.get_original_time``: a time on the concatenated (filtered) axis -> the original axis."""
    ends, silence = [], []
    previous_end = silent = 0
    for c in chunks:
        silent += c["start"] - previous_end
        previous_end = c["end"]
        ends.append(c["end"] - silent)
        silence.append(silent)
    sample = int(time * sampling_rate)
    idx = min(int(np.searchsorted(np.asarray(ends), sample, side="right")), len(ends) - 1)
    return round(silence[idx] / sampling_rate + time, time_precision)


def energy_probs(audio: np.ndarray, center_db: float = -40.0, width_db: float = 4.0) -> np.ndarray:
    """The stand-in speech-probability model the product ships until Silero weights exist (NOT upstream's network):
    per 512-sample window of the zero-padded signal, p = sigmoid((20 log10(rms + 1e-10) - center_db) / width_db)."""
    n = n_windows(audio.shape[0])
    x = np.zeros(n * WINDOW, np.float32)
    x[: audio.shape[0]] = audio
    rms = np.sqrt((x.reshape(n, WINDOW).astype(np.float64) ** 2).mean(axis=1))
    db = 20.0 * np.log10(rms + 1e-10)
    return (1.0 / (1.0 + np.exp(-(db - center_db) / width_db))).astype(np.float32)
