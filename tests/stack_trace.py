import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from gpu_diag_decode import fast_decoder_weights
from whisper_aries_b200 import WhisperDecoder, synthetic
B = int(sys.argv[1]); cta = sys.argv[2]
os.environ["ARIES_STACK_TRACE"] = cta
shape = synthetic.DEC_SHAPES["large-v3"]
tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
dec = WhisperDecoder(shape, fast_decoder_weights(shape), tokens=tok, max_batch=B)
enc = torch.randn(B, shape.n_audio_ctx, shape.d_model, device="cuda").bfloat16()
prompt = [tok.sot, tok.first_lang, tok.transcribe]
for _ in range(2):
    dec.generate(enc, [prompt] * B, max_length=3 + 32, suppress_tokens=[tok.eot])
print(dec.last_stats())
