"""Target for compute-sanitizer (run under gpurun):  compute-sanitizer --tool memcheck python tests/sanitize_target.py
Small decode runs that touch every row-f1 kernel variant: fused-LayerNorm and plain skinny GEMMs (all epilogues, cluster
split-K), self / cross / split attention, embedding, sampling with and without the timestamp rules."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_aries_b200 import WhisperDecoder, synthetic   # noqa: E402

os.environ["ARIES_DECODE_GRAPH"] = sys.argv[1] if len(sys.argv) > 1 else "0"
shape = synthetic.DEC_SHAPES["micro"]
tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
dec = WhisperDecoder(shape, synthetic.decoder_weights(shape, 32), tokens=tok, max_batch=16)
for batch, ts in ((3, True), (9, False), (16, True)):
    enc = torch.randn(batch, shape.n_audio_ctx, shape.d_model, device="cuda").bfloat16()
    prompt = [tok.sot, tok.first_lang, tok.transcribe] + ([] if ts else [tok.no_timestamps])
    res = dec.generate(enc, [prompt] * batch, max_length=len(prompt) + 6, suppress_tokens=[5, 6], return_scores=True,
                       return_no_speech_prob=True)
    print(batch, ts, [r.sequences_ids[0] for r in res][:2], dec.last_stats())
torch.cuda.synchronize()
print("ok")
