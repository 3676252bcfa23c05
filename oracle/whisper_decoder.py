"""fp32 torch-CPU restatement of the Whisper text decoder + greedy ``generate`` (oracle; test infrastructure) — row f1.

What the reference runs right after the encoder: ``model.transcribe(..., beam_size=1, best_of=1, temperature=0,
condition_on_previous_text=False)`` (ref: final_optimized_transcriber.py:326 with the knobs at :432-441) ->
faster-whisper ``generate_with_fallback`` -> ``ctranslate2.models.Whisper.generate(encoder_output, [prompt],
beam_size=1, max_length=448, suppress_blank=True, suppress_tokens=[...], max_initial_timestamp_index=50,
return_scores=True, return_no_speech_prob=True)`` -> ``layers::WhisperDecoder`` + ``GreedySearch`` with the logits
processors SuppressTokensBegin / SuppressTokens / ApplyTimestampRules (CT2 ``src/models/whisper.cc``)
[all upstream, unverified offline: restated from the published algorithm, identical to OpenAI whisper
``decoding.py`` for the timestamp rules].

PARITY UNPINNED against the true reference (no ctranslate2 offline); pinned against the independent HF
transformers 5.5 ``WhisperDecoder`` (logits) and ``WhisperTimeStampLogitsProcessor`` (rules) by
``oracle/make_golden.py`` -> ``tests/golden/decoder_golden.npz``.

Math per layer (pre-norm; same as HF modeling_whisper.py WhisperDecoderLayer):
  x += O_s(softmax_causal(q k^T) v),  q = (LN(x) Wq + bq) / sqrt(64), k without bias
  x += O_c(softmax(q' K_enc^T) V_enc), K_enc | V_enc = enc Wkv + b (key bias zero), computed once per window
  x += W2 gelu_erf(W1 LN(x) + b1) + b2
logits = LN_f(x) E^T with the output projection tied to the token embedding unless ``decoder/projection/weight``
is present.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5
NEG_INF = float("-inf")


@dataclass
class GenerateOptions:
    """The subset of ``ctranslate2.models.Whisper.generate`` arguments greedy decoding uses (upstream defaults)."""
    max_length: int = 448                       # total positions, prompt included
    suppress_blank: bool = True
    suppress_tokens: list = field(default_factory=list)      # explicit ids (upstream's -1 expands to the config's list)
    max_initial_timestamp_index: int = 50


def _t(w, key):
    v = w[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v)).to(torch.float32)


class Decoder:
    """Holds the weights as torch tensors; ``logits(tokens, enc)`` recomputes the whole prefix (no cache: clarity over speed)."""

    def __init__(self, w: dict, shape, round_weights_bf16: bool = False):
        self.shape = shape
        self.w = {}
        for k, v in w.items():
            if not k.startswith("decoder/"):
                continue
            t = _t(w, k)
            if round_weights_bf16 and t.ndim >= 2:
                t = t.to(torch.bfloat16).to(torch.float32)
            self.w[k] = t
        self._cross = None

    def _mha(self, q, k, v, causal: bool):
        B, Tq, d = q.shape
        h = self.shape.n_heads
        hd = d // h
        q = q.view(B, Tq, h, hd).transpose(1, 2) * hd ** -0.5
        k = k.view(B, -1, h, hd).transpose(1, 2)
        v = v.view(B, -1, h, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        if causal:
            s = s + torch.full((Tq, Tq), NEG_INF).triu(1)
        return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, Tq, d)

    @torch.no_grad()
    def cross_kv(self, enc: torch.Tensor):
        """enc f32 [B, 1500, d] -> per layer (K, V), each [B, 1500, d] (computed once per window, as CT2 caches it)."""
        out = []
        for i in range(self.shape.n_layers):
            p = f"decoder/layer_{i}/attention"
            kv = F.linear(enc, self.w[f"{p}/linear_1/weight"], self.w[f"{p}/linear_1/bias"])
            out.append(kv.split(self.shape.d_model, -1))
        return out

    @torch.no_grad()
    def logits(self, tokens: torch.Tensor, enc: torch.Tensor, cross=None, emulate_bf16: bool = False) -> torch.Tensor:
        """tokens int64 [B, T], enc f32 [B, 1500, d] -> f32 [B, T, vocab].

        ``emulate_bf16`` rounds every tensor the CUDA path stores (LayerNorm outputs, projections, attention outputs and
        the cached keys / values to bf16; the residual stream to f16) while keeping the arithmetic in f32: the error of
        THAT pipeline against the plain f32 one is the noise floor of the number format, against which the GPU tests
        judge the kernels' own error."""
        W, d = self.w, self.shape.d_model
        bf = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if emulate_bf16 else (lambda t: t)
        hf = (lambda t: t.to(torch.float16).to(torch.float32)) if emulate_bf16 else (lambda t: t)
        cross = cross if cross is not None else self.cross_kv(enc)
        cross = [(bf(k), bf(v)) for k, v in cross]
        x = hf(W["decoder/embeddings/weight"][tokens] + W["decoder/position_encodings/encodings"][: tokens.shape[1]])
        for i in range(self.shape.n_layers):
            p = f"decoder/layer_{i}"
            y = bf(F.layer_norm(x, (d,), W[f"{p}/self_attention/layer_norm/gamma"], W[f"{p}/self_attention/layer_norm/beta"], LN_EPS))
            q, k, v = bf(F.linear(y, W[f"{p}/self_attention/linear_0/weight"], W[f"{p}/self_attention/linear_0/bias"])).split(d, -1)
            x = hf(x + F.linear(bf(self._mha(q, k, v, True)), W[f"{p}/self_attention/linear_1/weight"], W[f"{p}/self_attention/linear_1/bias"]))
            y = bf(F.layer_norm(x, (d,), W[f"{p}/attention/layer_norm/gamma"], W[f"{p}/attention/layer_norm/beta"], LN_EPS))
            q = bf(F.linear(y, W[f"{p}/attention/linear_0/weight"], W[f"{p}/attention/linear_0/bias"]))
            x = hf(x + F.linear(bf(self._mha(q, cross[i][0], cross[i][1], False)), W[f"{p}/attention/linear_2/weight"], W[f"{p}/attention/linear_2/bias"]))
            y = bf(F.layer_norm(x, (d,), W[f"{p}/ffn/layer_norm/gamma"], W[f"{p}/ffn/layer_norm/beta"], LN_EPS))
            h = bf(F.gelu(F.linear(y, W[f"{p}/ffn/linear_0/weight"], W[f"{p}/ffn/linear_0/bias"])))
            x = hf(x + F.linear(h, W[f"{p}/ffn/linear_1/weight"], W[f"{p}/ffn/linear_1/bias"]))
        x = bf(F.layer_norm(x, (d,), W["decoder/layer_norm/gamma"], W["decoder/layer_norm/beta"], LN_EPS))
        proj = W.get("decoder/projection/weight", W["decoder/embeddings/weight"])
        return x @ proj.T


def apply_rules(logits: np.ndarray, seq: list, sample_begin: int, tok, opts: GenerateOptions, timestamps: bool):
    """Logits processors of one greedy step for ONE sequence.  ``logits`` f64/f32 [vocab] (copied), ``seq`` = every
    token so far (prompt included).  Returns (processed logits, rule margin) where the rule margin is how far the
    "timestamp mass vs best text token" decision (the only data-dependent rule) is from flipping (inf if not taken).

    1. SuppressTokensBegin: at the first sampled position, " " and EOT are forbidden (suppress_blank).
    2. SuppressTokens: the explicit id list.
    3. ApplyTimestampRules (only when the prompt carries no <|notimestamps|>): <|notimestamps|> forbidden; timestamps
       come in pairs (after an unpaired timestamp only timestamps/EOT ... after a pair only text); timestamps never
       decrease (and a segment-closing timestamp must advance); the first sampled token is a timestamp no later than
       max_initial_timestamp_index; and if the probability mass on timestamps exceeds the best text token's, text is
       forbidden."""
    lg = np.array(logits, dtype=np.float64)
    sampled = seq[sample_begin:]
    if opts.suppress_blank and len(sampled) == 0:
        lg[tok.blank] = NEG_INF
        lg[tok.eot] = NEG_INF
    for s in opts.suppress_tokens:
        lg[s] = NEG_INF
    margin = float("inf")
    if timestamps:
        tb = tok.timestamp_begin
        lg[tok.no_timestamps] = NEG_INF
        last_ts = len(sampled) >= 1 and sampled[-1] >= tb
        penult_ts = len(sampled) < 2 or sampled[-2] >= tb
        if last_ts:
            if penult_ts:
                lg[tb:] = NEG_INF
            else:
                lg[: tok.eot] = NEG_INF
        ts_seen = [t for t in sampled if t >= tb]
        if ts_seen:
            last = ts_seen[-1] if (last_ts and not penult_ts) else ts_seen[-1] + 1
            lg[tb:last] = NEG_INF
        if len(sampled) == 0:
            lg[:tb] = NEG_INF
            if opts.max_initial_timestamp_index is not None:
                lg[tb + opts.max_initial_timestamp_index + 1:] = NEG_INF
        m = lg.max()
        if np.isfinite(m):
            with np.errstate(divide="ignore"):
                ts_lse = np.log(np.exp(lg[tb:] - m).sum()) + m if np.isfinite(lg[tb:]).any() else NEG_INF
            text_max = lg[:tb].max()
            if np.isfinite(ts_lse) and np.isfinite(text_max):
                margin = abs(float(ts_lse - text_max))
            if ts_lse > text_max:
                lg[:tb] = NEG_INF
    return lg, margin


@torch.no_grad()
def generate(dec: Decoder, enc, prompts: list, tok, opts: GenerateOptions | None = None, forced: list | None = None):
    """Greedy decoding of every window of ``enc`` ([B, 1500, d], any float dtype).  ``prompts`` = one id list per window
    (all the same length, e.g. [sot, lang, transcribe]).  ``forced`` (tests only) replaces the argmax by the given
    continuation while still reporting what the argmax was (teacher forcing).

    Returns a list of dicts: sequences_ids (without prompt, without EOT), score (sum of log-probs of the processed
    distribution), no_speech_prob, argmax (token chosen at each step before forcing), margins (per step:
    min(top1 - top2 of the processed logits, timestamp-rule margin))."""
    opts = opts or GenerateOptions()
    enc = torch.as_tensor(np.asarray(enc.detach().to("cpu", torch.float32)) if isinstance(enc, torch.Tensor) else enc,
                          dtype=torch.float32)
    B = enc.shape[0]
    assert len(prompts) == B and len({len(p) for p in prompts}) == 1
    P = len(prompts[0])
    cross = dec.cross_kv(enc)
    seqs = [list(p) for p in prompts]
    res = [{"sequences_ids": [], "score": 0.0, "no_speech_prob": 0.0, "argmax": [], "margins": [], "done": False}
           for _ in range(B)]
    timestamps = [tok.no_timestamps not in p for p in prompts]
    sot_index = [p.index(tok.sot) if tok.sot in p else 0 for p in prompts]
    step = 0
    while len(seqs[0]) < opts.max_length and not all(r["done"] for r in res):
        lg_all = dec.logits(torch.tensor(seqs, dtype=torch.long), enc, cross)        # [B, T, V]
        for b in range(B):
            r = res[b]
            if step == 0:
                pr = torch.softmax(lg_all[b, sot_index[b]].double(), -1)
                r["no_speech_prob"] = float(pr[tok.no_speech])
            if r["done"]:
                seqs[b].append(tok.eot)
                continue
            lg, rule_margin = apply_rules(lg_all[b, -1].numpy(), seqs[b], P, tok, opts, timestamps[b])
            order = np.argsort(lg)[::-1]
            best = int(order[0])
            top_margin = float(lg[order[0]] - lg[order[1]]) if np.isfinite(lg[order[1]]) else float("inf")
            r["argmax"].append(best)
            r["margins"].append(min(top_margin, rule_margin))
            nxt = best
            if forced is not None and len(r["argmax"]) <= len(forced[b]):
                nxt = int(forced[b][len(r["argmax"]) - 1])
            m = lg.max()
            r["score"] += float(lg[nxt] - (np.log(np.exp(lg - m).sum()) + m))
            if nxt == tok.eot:
                r["done"] = True
            else:
                r["sequences_ids"].append(nxt)
            seqs[b].append(nxt)
        step += 1
    return res


@torch.no_grad()
def detect_language(dec: Decoder, enc, tok, lang_ids: list):
    """``ctranslate2.models.Whisper.detect_language`` [upstream, unverified offline; same as OpenAI whisper
    ``detect_language``]: logits of one decoder step on <|startoftranscript|>, everything but the language tokens
    masked, softmax.  Returns f64 [B, len(lang_ids)] in the order of ``lang_ids``."""
    enc = torch.as_tensor(enc, dtype=torch.float32)
    lg = dec.logits(torch.full((enc.shape[0], 1), tok.sot, dtype=torch.long), enc)[:, 0].double()
    return torch.softmax(lg[:, lang_ids], dim=-1).numpy()
