#!/usr/bin/env python
"""Writes the measured parity numbers north_star asks to be *stated* (run on a B200 under gpurun):

    python tests/gpu_parity_report.py [out.json]        # default gpurun_out/parity.json -> copied to profiles/rNN/

  * log-mel max-abs error vs the numpy oracle, 80 and 128 bins, the three signal families of SURVEY.md 8d;
  * encoder max-abs / cosine / worst-row cosine vs the fp32 torch oracle for tiny, medium and large-v3 (one window);
  * the BENCHMARKED launch: one B = 64 large-v3 ``encode_audio`` call (M = 96 000 rows) -- four scattered windows vs the
    oracle, and byte equality with the same windows encoded alone;
  * greedy probe token ids (oracle/decoder.py) from both outputs.
The oracle is only the checker here (test infrastructure)."""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import encoder as oenc, logmel as omel, synth as osynth
from oracle.decoder import GreedyProbe


def mel_report(n_mels: int) -> dict:
    from whisper_aries_b200 import FeatureExtractor
    fx = FeatureExtractor(feature_size=n_mels, device="cuda:0")
    worst, per = 0.0, {}
    for name, sig in (("tone_noise", osynth.tone_noise(0)), ("am_chirp", osynth.am_chirp(1)), ("gapped", osynth.gapped(2))):
        got = fx(sig)
        want = omel.log_mel(sig, n_mels)
        err = float(np.abs(got - want).max())
        per[name] = err
        worst = max(worst, err)
    return {"n_mels": n_mels, "max_abs": worst, "per_signal": per, "tolerance": 1e-4}


def encoder_report(name: str, batch64: bool) -> dict:
    from whisper_aries_b200 import WhisperModel, synthetic
    shape = synthetic.SHAPES[name]
    w = synthetic.encoder_weights(shape, 1234)
    model = WhisperModel(name, w, device="cuda", device_index=0)
    rep = {"shape": name}
    pcm1 = osynth.am_chirp(1)
    feats = omel.log_mel_window(pcm1, shape.n_mels)[None]
    t0 = time.perf_counter()
    ref = oenc.encoder_forward(feats, w, shape)
    rep["oracle_seconds_per_window"] = time.perf_counter() - t0
    out = model.encode(feats)
    rep["b1"] = oenc.compare(out.cpu(), ref)
    fused = model.encode_audio(torch.from_numpy(pcm1).cuda()[None])
    rep["b1_fused_pcm"] = oenc.compare(fused.cpu(), ref)
    probe, toks, margin = GreedyProbe.pick(ref, shape.d_model, shape.n_heads, steps=12, max_tries=64, n_layers=1)
    got, _ = probe.greedy(out.cpu(), steps=12)
    rep["probe_tokens"] = {"steps": 12, "margin": float(margin), "identical": bool(torch.equal(got, toks))}
    if batch64:
        B = 64
        pcm = osynth.batch_signals(B, 0)
        dev = torch.from_numpy(pcm).cuda()
        big = model.encode_audio(dev)                       # ONE launch sequence, M = 96 000 rows
        torch.cuda.synchronize()
        rep["b64"] = {"finite": bool(torch.isfinite(big.float()).all()), "windows": {}}
        worst = {"max_abs": 0.0, "cosine": 1.0, "min_row_cosine": 1.0}
        for k in (0, 21, 42, 63):
            alone = model.encode_audio(dev[k:k + 1])
            f = omel.log_mel_window(pcm[k], shape.n_mels)[None]
            c = oenc.compare(big[k:k + 1].cpu(), oenc.encoder_forward(f, w, shape))
            c["bytes_equal_to_b1"] = bool(torch.equal(alone[0], big[k]))
            rep["b64"]["windows"][str(k)] = c
            worst["max_abs"] = max(worst["max_abs"], c["max_abs"])
            worst["cosine"] = min(worst["cosine"], c["cosine"])
            worst["min_row_cosine"] = min(worst["min_row_cosine"], c["min_row_cosine"])
        rep["b64"]["worst"] = worst
        rep["b64"]["all_bytes_equal_to_b1"] = all(v["bytes_equal_to_b1"] for v in rep["b64"]["windows"].values())
    del model
    torch.cuda.empty_cache()
    return rep


def main() -> int:
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity.json")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    rep = {"device": torch.cuda.get_device_name(0), "host_cores": os.cpu_count(),
           "oracle": "oracle/logmel.py + oracle/encoder.py (fp32; parity unpinned against faster-whisper/CTranslate2: "
                     "not installable offline)",
           "mel": [mel_report(80), mel_report(128)],
           "encoder": [encoder_report("tiny", False), encoder_report("medium", False), encoder_report("large-v3", True)]}
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))
    return 0


if __name__ == "__main__":
    sys.exit(main())
