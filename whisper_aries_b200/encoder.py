"""``encode`` half of ``faster_whisper.WhisperModel`` / ``ctranslate2.models.Whisper`` on a B200.

Upstream: ``WhisperModel.encode(features)`` -> ``ctranslate2.models.Whisper.encode`` -> ``layers::WhisperEncoder``
(SURVEY.md rows a-5..a-8), reached by the reference through ``model.transcribe`` (ref:
final_optimized_transcriber.py:326; model built at :179-182 and forced to large-v3 at conversation_transcriber.py:72).
Here the forward pass is hand-written sm_100a CUDA (tcgen05 GEMMs with fused epilogues, fused attention, LayerNorm)
behind ``aries_encoder_run``; this file only moves pointers."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .feature_extractor import FeatureExtractor
from .synthetic import SHAPES, EncoderShape


class WhisperEncoder:
    """Owns one ``aries_encoder`` handle (weights resident on one GPU) plus a growable device workspace."""

    def __init__(self, shape: EncoderShape | str, weights: dict, device="cuda:0"):
        import torch
        self.shape = SHAPES[shape] if isinstance(shape, str) else shape
        self.device_index = _lib.device_index_of(device)
        self.device = torch.device("cuda", self.device_index)
        self._ctx = _lib.Context.get(self.device_index)
        lib = self._ctx.lib
        cfg = _lib.EncoderCfg(self.shape.n_mels, self.shape.d_model, self.shape.n_heads, self.shape.n_layers,
                              self.shape.d_ffn, self.shape.n_ctx)
        keep, descs = [], (_lib.WeightDesc * len(weights))()
        for i, (name, arr) in enumerate(weights.items()):
            a = np.ascontiguousarray(np.asarray(arr), dtype=np.float32)
            if a.ndim < 1 or a.ndim > 4:
                raise ValueError(f"weight {name!r} has unsupported rank {a.ndim}")
            keep.append(a)
            descs[i].name = name.encode()
            descs[i].data = a.ctypes.data
            descs[i].ndim = a.ndim
            for k in range(4):
                descs[i].shape[k] = a.shape[k] if k < a.ndim else 1
        h = ctypes.c_void_p()
        _lib.check(lib.aries_encoder_create(self._ctx.handle, ctypes.byref(cfg), descs, len(weights), ctypes.byref(h)))
        self._handle = h
        self._ws = None
        self._ws_batch = 0

    def close(self):
        if getattr(self, "_handle", None) is not None:
            self._ctx.lib.aries_encoder_destroy(self._handle)
            self._handle = None
            self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launches(self) -> int:
        return self._ctx.lib.aries_encoder_last_launches(self._handle)

    def set_profiling(self, on: bool) -> None:
        """Bracket every kernel of the following calls with CUDA events (bench.py's per-kernel roofline numbers)."""
        _lib.check(self._ctx.lib.aries_encoder_set_profiling(self._handle, int(bool(on))))

    def collect_profile(self) -> dict:
        """{kernel class: (milliseconds, launches)} accumulated since profiling was switched on / last collected."""
        n = len(_lib.KERNEL_CLASSES)
        ms, cnt = (ctypes.c_float * n)(), (ctypes.c_int * n)()
        _lib.check(self._ctx.lib.aries_encoder_collect_profile(self._handle, ms, cnt, n))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.KERNEL_CLASSES)}

    def workspace_bytes(self, batch: int) -> int:
        return int(self._ctx.lib.aries_encoder_workspace_bytes(self._handle, batch))

    def _workspace(self, batch: int):
        import torch
        if self._ws is None or batch > self._ws_batch:
            self._ws = None
            self._ws = torch.empty(self.workspace_bytes(batch), dtype=torch.uint8, device=self.device)
            self._ws_batch = batch
        return self._ws

    def _check(self, shape) -> None:
        if len(shape) != 3 or shape[1] != self.shape.n_mels or shape[2] > 3000 or shape[2] == 0 or shape[0] == 0:
            raise ValueError(f"Invalid input features shape: expected an input with shape (batch, {self.shape.n_mels}, "
                             f"<=3000), but got {tuple(shape)}")

    def encode(self, features):
        """features f32 ``[n_mels, frames]`` / ``[B, n_mels, frames]`` (numpy or torch, host or device), frames <= 3000
        -> CUDA bf16 tensor ``[B, 1500, d_model]`` (device-resident, like upstream's StorageView)."""
        import torch
        if isinstance(features, torch.Tensor):
            x = features
        else:
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(features), dtype=np.float32))
        if x.dim() == 2:
            x = x[None]
        self._check(tuple(x.shape))
        x = x.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        batch, _, frames = x.shape
        ws = self._workspace(batch)
        out = torch.empty((batch, self.shape.n_ctx, self.shape.d_model), dtype=torch.bfloat16, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._ctx.lib.aries_encoder_run(self._handle, x.data_ptr(), batch, frames, out.data_ptr(),
                                                   ws.data_ptr(), ws.numel(), stream))
        return out

    def encode_pcm(self, extractor: FeatureExtractor, pcm, out=None):
        """Fused PCM -> log-mel -> encoder for a CUDA f32 tensor ``[B, n_samples <= 480000]``; the mel tensor never
        leaves the GPU.  Returns CUDA bf16 ``[B, 1500, d_model]``."""
        import torch
        if not (isinstance(pcm, torch.Tensor) and pcm.is_cuda and pcm.dtype == torch.float32):
            raise ValueError("encode_pcm expects a CUDA float32 tensor [batch, n_samples]")
        if pcm.dim() == 1:
            pcm = pcm[None]
        if pcm.stride(1) != 1:
            pcm = pcm.contiguous()
        batch, n = pcm.shape
        if pcm.device != self.device:
            raise ValueError(f"pcm is on {pcm.device}, this encoder on {self.device}")
        ws = self._workspace(batch)
        want = (batch, self.shape.n_ctx, self.shape.d_model)
        if out is None:
            out = torch.empty(want, dtype=torch.bfloat16, device=self.device)
        elif not (isinstance(out, torch.Tensor) and out.is_cuda and out.device == self.device and
                  out.dtype == torch.bfloat16 and tuple(out.shape) == want and out.is_contiguous()):
            raise ValueError(f"out must be a contiguous CUDA bfloat16 tensor {want} on {self.device}")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._ctx.lib.aries_encode_pcm(self._handle, extractor._mel(), pcm.data_ptr(), batch, n,
                                                  pcm.stride(0), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return out


class WhisperModel:
    """The slice of ``faster_whisper.WhisperModel`` on the hot path: ``feature_extractor`` + ``encode``.

    ``WhisperModel("large-v3", weights=..., device="cuda", device_index=0)`` mirrors the reference's construction
    (ref: final_optimized_transcriber.py:179-182).  ``compute_type`` is accepted for signature compatibility; the
    B200 path always computes in bf16 with f32 accumulation.  When ``weights`` also holds the ``decoder/...`` variables
    and ``decoder_shape`` is given, ``generate`` (greedy ``ctranslate2.models.Whisper.generate``, row f1) is available
    too; tokenisation / segment assembly (``transcribe``) stay outside this library."""

    def __init__(self, model_size_or_shape, weights: dict | None = None, device: str = "cuda", device_index: int = 0,
                 compute_type: str = "bfloat16", decoder_shape=None, max_batch: int = 64, suppress_ids=(), **_ignored):
        if device not in ("cuda", "auto"):
            raise ValueError("whisper_aries_b200 has no CPU path: device must be 'cuda'")
        dev = f"cuda:{device_index}"
        self.model_info = None
        if weights is None:
            # upstream's first argument is "size or path of a converted model directory" (model.bin + config.json);
            # there is no network here, so only the path form can supply weights
            import os
            from .ct2_model import load_encoder_weights, parse_model_dir
            if not (isinstance(model_size_or_shape, str) and os.path.isdir(model_size_or_shape)):
                raise ValueError("weights=None needs the path of a CTranslate2 model directory (model.bin); "
                                 f"got {model_size_or_shape!r}")
            model_dir = model_size_or_shape
            parsed = parse_model_dir(model_dir)                   # model.bin is mapped and parsed once for both halves
            model_size_or_shape, weights, self.model_info = load_encoder_weights(model_dir, parsed)
            if decoder_shape is None:
                # the same model.bin carries the decoder: build it too, as upstream's WhisperModel(path) does
                from .ct2_model import load_decoder_weights
                try:
                    decoder_shape, dec_w, dec_info = load_decoder_weights(model_dir, parsed)
                except (KeyError, ValueError):
                    decoder_shape = None                       # an encoder-only directory
                else:
                    weights = dict(weights)
                    weights.update(dec_w)
                    suppress_ids = suppress_ids or dec_info["suppress_ids"]
        shape = SHAPES[model_size_or_shape] if isinstance(model_size_or_shape, str) else model_size_or_shape
        self.shape = shape
        self.compute_type = compute_type
        self.feature_extractor = FeatureExtractor(feature_size=shape.n_mels, device=dev)
        self.encoder = WhisperEncoder(shape, weights, device=dev)
        self.decoder = None
        if decoder_shape is not None:
            from .decoder import WhisperDecoder
            self.decoder = WhisperDecoder(decoder_shape, weights, device=dev, max_batch=max_batch,
                                          suppress_ids=suppress_ids)

    def generate(self, encoder_output, prompts, **kw):
        """``ctranslate2.models.Whisper.generate`` (greedy) on the encoder output of ``encode`` / ``encode_audio``."""
        if self.decoder is None:
            raise RuntimeError("this WhisperModel was built without decoder weights (pass decoder_shape=...)")
        return self.decoder.generate(encoder_output, prompts, **kw)

    def detect_language(self, encoder_output, lang_ids=None):
        """``ctranslate2.models.Whisper.detect_language`` on the encoder output (language=None path of ``transcribe``)."""
        if self.decoder is None:
            raise RuntimeError("this WhisperModel was built without decoder weights (pass decoder_shape=...)")
        return self.decoder.detect_language(encoder_output, lang_ids)

    def encode(self, features):
        return self.encoder.encode(features)

    def encode_audio(self, pcm, out=None):
        """PCM windows (CUDA f32, or int16 as decoded by ffmpeg, ``[B, <=480000]``) -> encoder states, fused on the
        device.  int16 input is converted on the GPU (``/ 32768`` as faster-whisper's ``decode_audio`` does on the host)."""
        import torch
        if isinstance(pcm, torch.Tensor) and pcm.dtype == torch.int16:
            pcm = pcm_s16_to_f32(pcm)
        return self.encoder.encode_pcm(self.feature_extractor, pcm, out=out)

    def encode_long(self, pcm, vad_filter: bool = False, vad_parameters=None, speech_probs=None, vad_model=None):
        """One waveform longer than 30 s, exactly as upstream's ``transcribe`` feeds it to the encoder when it steps
        ``seek`` by full windows: features over the WHOLE call (one clamp maximum, STFT frames that straddle a 30-s
        boundary see real samples on both sides, ref call site final_optimized_transcriber.py:326 with 185-s chunks),
        ``content_frames = frames - 1``, then ``pad_or_trim(features[:, seek:seek + 3000])`` per window -> CUDA bf16
        ``[ceil(content_frames / 3000), 1500, d_model]``.  ``pcm``: 1-D float32 (numpy or torch, host or device).

        ``vad_filter=True`` (what the reference always passes, ref: final_optimized_transcriber.py:440) first removes
        non-speech as upstream does -- ``get_speech_timestamps`` + ``collect_chunks`` on the device (``vad.py``; the
        probability model is pluggable, see there) -- and returns ``(states, speech_chunks)`` so that the caller can
        map times back with ``vad.SpeechTimestampsMap``."""
        import torch
        x = pcm if isinstance(pcm, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pcm, dtype=np.float32))
        if x.dim() != 1 or x.numel() == 0:
            raise ValueError("encode_long expects a non-empty 1-D waveform")
        x = x.to(device=self.encoder.device, dtype=torch.float32)
        chunks = None
        if vad_filter:
            from . import vad
            opts = vad_parameters if isinstance(vad_parameters, vad.VadOptions) else vad.VadOptions(**(vad_parameters or {}))
            chunks = vad.get_speech_timestamps(x, opts, speech_probs=speech_probs, model=vad_model)
            x = vad.collect_chunks(x, chunks)
            if x.numel() == 0:                     # no speech at all: nothing reaches the feature extractor
                return torch.empty((0, self.shape.n_ctx, self.shape.d_model), dtype=torch.bfloat16,
                                   device=self.encoder.device), chunks
        feats = self.feature_extractor(x)                                   # [n_mels, (N + 160) // 160], device
        content = feats.shape[-1] - 1
        n_win = max(1, -(-content // 3000))
        batch = torch.zeros((n_win, feats.shape[0], 3000), dtype=torch.float32, device=feats.device)
        for k in range(n_win):
            seg = feats[:, k * 3000: min((k + 1) * 3000, content)]
            batch[k, :, : seg.shape[1]] = seg
        out = self.encoder.encode(batch)
        return (out, chunks) if vad_filter else out


def pcm_s16_to_f32(pcm_s16, out=None):
    """CUDA int16 tensor -> CUDA float32 tensor of the same shape, ``x / 32768`` (``aries_pcm_s16_to_f32``)."""
    import torch
    if not (isinstance(pcm_s16, torch.Tensor) and pcm_s16.is_cuda and pcm_s16.dtype == torch.int16):
        raise ValueError("pcm_s16_to_f32 expects a CUDA int16 tensor")
    x = pcm_s16.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    elif out.shape != x.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("out must be a contiguous CUDA float32 tensor of the input's shape")
    ctx = _lib.Context.get(x.device.index or 0)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(ctx.lib.aries_pcm_s16_to_f32(ctx.handle, x.data_ptr(), out.data_ptr(), x.numel(), stream))
    return out
