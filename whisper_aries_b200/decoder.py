"""``generate`` half of ``ctranslate2.models.Whisper`` on a B200.

Upstream: faster-whisper's ``generate_with_fallback`` calls ``self.model.generate(encoder_output, [prompt],
beam_size=..., max_length=448, return_scores=True, return_no_speech_prob=True, suppress_blank=True,
suppress_tokens=[...], max_initial_timestamp_index=50)``; the reference pins ``beam_size=1, best_of=1, temperature=0,
condition_on_previous_text=False`` (ref: final_optimized_transcriber.py:432-441, reached from ``model.transcribe`` at
:326).  Here the decoder forward, the key/value caches, the logits rules and the argmax run as hand-written sm_100a
CUDA behind ``aries_decoder_generate`` (one CUDA graph per token); this file only moves pointers and mirrors the
upstream call surface.  Token ids only: the tokenizer is outside the hot path (the caller owns it, as upstream)."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .synthetic import DEC_SHAPES, DecoderShape, WhisperTokens


@dataclass
class WhisperGenerationResult:
    """Mirror of ``ctranslate2.models.WhisperGenerationResult`` (greedy: one hypothesis)."""
    sequences_ids: list
    scores: list
    no_speech_prob: float
    sequences: list = field(default_factory=list)      # token strings need the tokenizer: not produced here


class WhisperDecoder:
    """Owns one ``aries_decoder`` handle: decoder weights, self-attention cache for ``max_batch`` sequences, and the
    cross-attention key/value cache of the batch being decoded (all resident on one GPU)."""

    def __init__(self, shape: DecoderShape | str, weights: dict, tokens: WhisperTokens | None = None, device="cuda:0",
                 max_batch: int = 64, suppress_ids=()):
        import torch
        self.shape = DEC_SHAPES[shape] if isinstance(shape, str) else shape
        self.tokens = tokens or WhisperTokens.for_vocab(self.shape.vocab)
        self.suppress_ids = [int(t) for t in suppress_ids]
        self.max_batch = int(max_batch)
        self._warned_no_suppress = False
        self.device_index = _lib.device_index_of(device)
        self.device = torch.device("cuda", self.device_index)
        self._ctx = _lib.Context.get(self.device_index)
        cfg = _lib.DecoderCfg(self.shape.vocab, self.shape.d_model, self.shape.n_heads, self.shape.n_layers,
                              self.shape.d_ffn, self.shape.n_text_ctx, self.shape.n_audio_ctx)
        names = [k for k in weights if k.startswith("decoder/")]
        keep, descs = [], (_lib.WeightDesc * len(names))()
        for i, name in enumerate(names):
            a = np.ascontiguousarray(np.asarray(weights[name]), dtype=np.float32)
            keep.append(a)
            descs[i].name = name.encode()
            descs[i].data = a.ctypes.data
            descs[i].ndim = a.ndim
            for k in range(4):
                descs[i].shape[k] = a.shape[k] if k < a.ndim else 1
        h = ctypes.c_void_p()
        _lib.check(self._ctx.lib.aries_decoder_create(self._ctx.handle, ctypes.byref(cfg), descs, len(names),
                                                      self.max_batch, ctypes.byref(h)))
        self._handle = h

    def close(self):
        if getattr(self, "_handle", None) is not None:
            self._ctx.lib.aries_decoder_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_stats(self) -> dict:
        out = (ctypes.c_float * 5)()
        _lib.check(self._ctx.lib.aries_decoder_last_stats(self._handle, out, 5))
        return {"cross_kv_ms": out[0], "decode_ms": out[1], "steps": int(out[2]), "kernels_per_step": int(out[3]),
                "cross_kv_kernels": int(out[4])}

    # ------------------------------------------------------------------------------------------------------------
    def _opts(self, max_length, suppress_blank, suppress_tokens, max_initial_timestamp_index):
        t = self.tokens
        ids = []
        for s in suppress_tokens or ():
            if int(s) == -1:
                if not self.suppress_ids and not self._warned_no_suppress:
                    # upstream reads them from the model's config.json; a decoder built from a weights dict has none
                    import warnings
                    warnings.warn("suppress_tokens=-1 requested but this decoder has no suppress_ids (pass "
                                  "suppress_ids=... or load a model directory with config.json): decoding WITHOUT "
                                  "non-speech token suppression", RuntimeWarning, stacklevel=3)
                    self._warned_no_suppress = True
                ids.extend(self.suppress_ids)        # upstream: -1 expands to the model config's suppress_ids
            else:
                ids.append(int(s))
        ids = sorted(set(ids))
        arr = (ctypes.c_int32 * max(1, len(ids)))(*ids)
        o = _lib.GenerateOpts(int(max_length), int(bool(suppress_blank)), t.blank, t.eot, t.sot, t.no_speech,
                              t.no_timestamps, t.timestamp_begin, int(max_initial_timestamp_index),
                              ctypes.cast(arr, ctypes.POINTER(ctypes.c_int32)), len(ids))
        return o, arr

    def _check_inputs(self, encoder_output, prompts):
        import torch
        if not (isinstance(encoder_output, torch.Tensor) and encoder_output.is_cuda and
                encoder_output.dtype == torch.bfloat16):
            raise ValueError("generate expects the encoder output as a CUDA bfloat16 tensor [batch, 1500, d_model] "
                             "(what WhisperModel.encode returns)")
        if encoder_output.dim() != 3 or encoder_output.shape[1] != self.shape.n_audio_ctx or \
                encoder_output.shape[2] != self.shape.d_model:
            raise ValueError(f"Invalid encoder output shape: expected (batch, {self.shape.n_audio_ctx}, "
                             f"{self.shape.d_model}), but got {tuple(encoder_output.shape)}")
        if len(prompts) != encoder_output.shape[0]:
            raise ValueError("generate needs one prompt per window")
        if any(not isinstance(t, (int, np.integer)) for p in prompts for t in p):
            raise ValueError("prompts must be token ids (the tokenizer is outside this library)")
        if len({len(p) for p in prompts}) != 1 or len(prompts[0]) == 0:
            raise ValueError("all prompts of a batch must have the same, non-zero length")

    def language_ids(self) -> list:
        """Ids of the language tokens: everything between <|startoftranscript|>'s successor and <|translate|>."""
        t = self.tokens
        return list(range(t.first_lang, min(t.translate, t.transcribe)))

    def detect_language(self, encoder_output, lang_ids=None):
        """``ctranslate2.models.Whisper.detect_language``: for every window the language tokens with their probability,
        most probable first -- upstream returns ``("<|en|>", p)`` pairs, here ``(token id, p)`` (no tokenizer in this
        library).  faster-whisper calls it when ``language=None``, the reference's default (ref:
        final_optimized_transcriber.py:433, recorded at :350-351)."""
        import torch
        self._check_inputs(encoder_output, [[self.tokens.sot]] * encoder_output.shape[0])
        ids = [int(i) for i in (lang_ids if lang_ids is not None else self.language_ids())]
        if not ids:
            raise ValueError("no language token ids")
        arr = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
        opts, _keep = self._opts(2, False, (), 50)
        enc = encoder_output.contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        out = []
        for b0 in range(0, enc.shape[0], self.max_batch):
            part = enc[b0:b0 + self.max_batch]
            probs = np.zeros((part.shape[0], len(ids)), dtype=np.float32)
            _lib.check(self._ctx.lib.aries_decoder_detect_language(self._handle, part.data_ptr(), part.shape[0],
                                                                   ctypes.byref(opts), arr.ctypes.data, len(ids),
                                                                   probs.ctypes.data, stream))
            for row in probs:
                order = np.argsort(-row, kind="stable")
                out.append([(ids[i], float(row[i])) for i in order])
        return out

    def generate(self, encoder_output, prompts, *, beam_size: int = 1, patience: float = 1, num_hypotheses: int = 1,
                 length_penalty: float = 1, repetition_penalty: float = 1, no_repeat_ngram_size: int = 0,
                 max_length: int = 448, return_scores: bool = False, return_no_speech_prob: bool = False,
                 max_initial_timestamp_index: int = 50, suppress_blank: bool = True, suppress_tokens=(-1,),
                 sampling_topk: int = 1, sampling_temperature: float = 1, _forced=None, _want_logits: bool = False):
        """``ctranslate2.models.Whisper.generate`` for greedy decoding -> list of ``WhisperGenerationResult``.

        ``encoder_output``: CUDA bf16 ``[B, 1500, d]`` from ``WhisperModel.encode``; ``prompts``: one id list per window.
        ``max_length`` counts every position, prompt included.  Anything but beam_size=1 / sampling_topk=1 (what the
        reference uses) raises ValueError."""
        import torch
        if beam_size != 1 or num_hypotheses != 1 or sampling_topk != 1:
            raise ValueError("whisper_aries_b200 decodes greedily (beam_size=1, sampling_topk=1), as the reference does")
        if repetition_penalty != 1 or no_repeat_ngram_size != 0:
            raise ValueError("repetition_penalty / no_repeat_ngram_size are not supported")
        self._check_inputs(encoder_output, prompts)
        max_length = min(int(max_length), self.shape.n_text_ctx)
        P = len(prompts[0])
        if P >= max_length:
            raise ValueError("the prompt leaves no room to generate (prompt length >= max_length)")
        opts, _keep = self._opts(max_length, suppress_blank, suppress_tokens, max_initial_timestamp_index)
        enc = encoder_output.contiguous()
        lib = self._ctx.lib
        stream = torch.cuda.current_stream(self.device).cuda_stream
        results, extras = [], []
        for b0 in range(0, enc.shape[0], self.max_batch):
            part = enc[b0:b0 + self.max_batch]
            B = part.shape[0]
            pr = np.ascontiguousarray(np.asarray(prompts[b0:b0 + B], dtype=np.int32))
            if _forced is None and not _want_logits:
                toks, lens, scores, nsp = self._generate_arrays(part, pr, opts, max_length, stream)
            else:
                toks = np.empty((B, max_length), dtype=np.int32)
                lens = np.zeros(B, dtype=np.int32)
                scores = np.zeros(B, dtype=np.float32)
                nsp = np.zeros(B, dtype=np.float32)
                forced = np.ascontiguousarray(np.asarray(_forced[b0:b0 + B], dtype=np.int32)) if _forced is not None \
                    else np.zeros((B, 0), dtype=np.int32)
                argmax = np.full((B, max_length), -1, dtype=np.int32)
                logits = np.zeros((max_length - 1, B, self.shape.vocab), dtype=np.float32) if _want_logits else None
                _lib.check(lib.aries_test_decoder_generate(
                    self._handle, part.data_ptr(), B, pr.ctypes.data, P, ctypes.byref(opts),
                    forced.ctypes.data if forced.size else None, forced.shape[1], toks.ctypes.data, argmax.ctypes.data,
                    logits.ctypes.data if logits is not None else None, scores.ctypes.data, nsp.ctypes.data, stream))
                eot = self.tokens.eot
                for b in range(B):
                    n = 0
                    for t in toks[b, P:]:
                        if t == eot:
                            break
                        n += 1
                    lens[b] = n
                extras.append({"argmax": argmax, "logits": logits, "tokens": toks.copy()})
            for b in range(B):
                ids = toks[b, P:P + lens[b]].tolist()
                # upstream normalises the cumulative log-prob (EOT's included) by length ** length_penalty
                # [unverified offline]; faster-whisper multiplies it back (cum_logprob = score * seq_len ** penalty)
                norm = float(max(len(ids), 1)) ** float(length_penalty)
                results.append(WhisperGenerationResult([ids], [float(scores[b]) / norm] if return_scores else [],
                                                       float(nsp[b]) if return_no_speech_prob else 0.0))
        if _forced is not None or _want_logits:
            return results, extras
        return results

    def _generate_arrays(self, part, prompts_i32, opts, max_length: int, stream):
        """One ``aries_decoder_generate`` call for <= max_batch windows -> (tokens [B, max_length] int32 = prompt, sampled
        ids, EOT padding; sampled-id counts [B]; cumulative log-probs [B]; no-speech probabilities [B])."""
        B, P = prompts_i32.shape
        toks = np.empty((B, max_length), dtype=np.int32)
        lens = np.zeros(B, dtype=np.int32)
        scores = np.zeros(B, dtype=np.float32)
        nsp = np.zeros(B, dtype=np.float32)
        _lib.check(self._ctx.lib.aries_decoder_generate(self._handle, part.data_ptr(), B, prompts_i32.ctypes.data, P,
                                                        ctypes.byref(opts), toks.ctypes.data, lens.ctypes.data,
                                                        scores.ctypes.data, nsp.ctypes.data, stream))
        return toks, lens, scores, nsp

    def generate_ids(self, encoder_output, prompt, *, max_length: int = 448, max_initial_timestamp_index: int = 50,
                     suppress_blank: bool = True, suppress_tokens=(-1,)):
        """Batch form used by ``gpu_transcribe_worker``: the same greedy ``generate`` with ONE prompt for every window,
        returning arrays instead of per-window result objects -> (tokens int32 ``[B, max_length]``, counts ``[B]``)."""
        import torch
        B = encoder_output.shape[0]
        self._check_inputs(encoder_output, [list(prompt)] * B)
        max_length = min(int(max_length), self.shape.n_text_ctx)
        if len(prompt) >= max_length:
            raise ValueError("the prompt leaves no room to generate (prompt length >= max_length)")
        opts, _keep = self._opts(max_length, suppress_blank, suppress_tokens, max_initial_timestamp_index)
        enc = encoder_output.contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        toks, lens = [], []
        for b0 in range(0, B, self.max_batch):
            part = enc[b0:b0 + self.max_batch]
            pr = np.ascontiguousarray(np.tile(np.asarray(prompt, dtype=np.int32), (part.shape[0], 1)))
            t, n, _, _ = self._generate_arrays(part, pr, opts, max_length, stream)
            toks.append(t)
            lens.append(n)
        return np.concatenate(toks), np.concatenate(lens)
