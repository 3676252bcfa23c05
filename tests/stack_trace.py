"""Per-phase clock64 timeline of the opt-in persistent decode kernel (csrc/decode_stack.cu), run under gpurun:

    python tests/stack_trace.py <windows> <cta>

Sets ARIES_DECODE_STACK=1 and ARIES_STACK_TRACE=<cta>; the library prints, for the first three layers of the last token
step, the cycles between consecutive stamps of that CTA's thread 0 -- per layer: LN1, QKV, barrier | self-attention,
barrier | context staging, O, barrier | LN2, Q, barrier | cross-attention, barrier | context staging, O, barrier | LN3,
fc1, barrier | fc1-output staging, fc2, barrier.  The numbers quoted in DESIGN.md (row f1) come from this."""
import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from gpu_diag_decode import fast_decoder_weights
from whisper_aries_b200 import WhisperDecoder, synthetic
B = int(sys.argv[1]); cta = sys.argv[2]
os.environ["ARIES_STACK_TRACE"] = cta
os.environ["ARIES_DECODE_STACK"] = "1"
shape = synthetic.DEC_SHAPES["large-v3"]
tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
dec = WhisperDecoder(shape, fast_decoder_weights(shape), tokens=tok, max_batch=B)
enc = torch.randn(B, shape.n_audio_ctx, shape.d_model, device="cuda").bfloat16()
prompt = [tok.sot, tok.first_lang, tok.transcribe]
for _ in range(2):
    dec.generate(enc, [prompt] * B, max_length=3 + 32, suppress_tokens=[tok.eot])
print(dec.last_stats())
