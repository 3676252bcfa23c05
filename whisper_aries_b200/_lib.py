"""ctypes binding of libaries_b200.so (include/aries_b200.h).

There is no CPU fallback: importing this module without the built library, or using it on a machine without an
sm_100 GPU, raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C
whisper_aries_b200/csrc``)."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libaries_b200.so")

ARIES_OK, ARIES_EINVAL, ARIES_ECUDA, ARIES_ENOMEM, ARIES_ESTATE = 0, -1, -2, -3, -4

c_void_p, c_int, c_int64, c_size_t, c_char_p = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t,
                                                ctypes.c_char_p)
c_float_p = ctypes.POINTER(ctypes.c_float)


class EncoderCfg(ctypes.Structure):
    _fields_ = [("n_mels", ctypes.c_int32), ("d_model", ctypes.c_int32), ("n_heads", ctypes.c_int32),
                ("n_layers", ctypes.c_int32), ("d_ffn", ctypes.c_int32), ("n_ctx", ctypes.c_int32)]


class WeightDesc(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("data", c_void_p), ("ndim", ctypes.c_int32), ("shape", ctypes.c_int64 * 4)]


class DecoderCfg(ctypes.Structure):
    _fields_ = [("vocab", ctypes.c_int32), ("d_model", ctypes.c_int32), ("n_heads", ctypes.c_int32),
                ("n_layers", ctypes.c_int32), ("d_ffn", ctypes.c_int32), ("n_text_ctx", ctypes.c_int32),
                ("n_audio_ctx", ctypes.c_int32)]


class GenerateOpts(ctypes.Structure):
    _fields_ = [("max_length", ctypes.c_int32), ("suppress_blank", ctypes.c_int32), ("blank_id", ctypes.c_int32),
                ("eot", ctypes.c_int32), ("sot", ctypes.c_int32), ("no_speech", ctypes.c_int32),
                ("no_timestamps", ctypes.c_int32), ("timestamp_begin", ctypes.c_int32),
                ("max_initial_timestamp_index", ctypes.c_int32), ("suppress_tokens", ctypes.POINTER(ctypes.c_int32)),
                ("n_suppress", ctypes.c_int32)]


class VadOpts(ctypes.Structure):
    _fields_ = [("threshold", ctypes.c_float), ("neg_threshold", ctypes.c_float),
                ("min_speech_duration_ms", ctypes.c_int32), ("max_speech_duration_s", ctypes.c_float),
                ("min_silence_duration_ms", ctypes.c_int32), ("speech_pad_ms", ctypes.c_int32)]


c_float = ctypes.c_float

# name -> (restype, argtypes); every symbol include/aries_b200.h and include/aries_b200_test.h declare
PROTOTYPES = {
    "aries_abi_version": (c_int, []),
    "aries_last_error": (c_char_p, []),
    "aries_init": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "aries_destroy": (c_int, [c_void_p]),
    "aries_device": (c_int, [c_void_p]),
    "aries_sm_count": (c_int, [c_void_p]),
    "aries_logmel_create": (c_int, [c_void_p, c_int, c_void_p, ctypes.POINTER(c_void_p)]),
    "aries_logmel_destroy": (c_int, [c_void_p]),
    "aries_logmel_num_frames": (c_int64, [c_int64, c_int]),
    "aries_logmel_run": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "aries_logmel_run_host": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int]),
    "aries_logmel_last_launches": (c_int, [c_void_p]),
    "aries_encoder_create": (c_int, [c_void_p, ctypes.POINTER(EncoderCfg), ctypes.POINTER(WeightDesc), c_int,
                                     ctypes.POINTER(c_void_p)]),
    "aries_encoder_destroy": (c_int, [c_void_p]),
    "aries_encoder_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "aries_encoder_run": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aries_encoder_run_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "aries_encode_pcm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "aries_encoder_last_launches": (c_int, [c_void_p]),
    "aries_pcm_s16_to_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "aries_encoder_set_profiling": (c_int, [c_void_p, c_int]),
    "aries_encoder_collect_profile": (c_int, [c_void_p, c_float_p, ctypes.POINTER(c_int), c_int]),
    "aries_decoder_create": (c_int, [c_void_p, ctypes.POINTER(DecoderCfg), ctypes.POINTER(WeightDesc), c_int, c_int,
                                     ctypes.POINTER(c_void_p)]),
    "aries_decoder_destroy": (c_int, [c_void_p]),
    "aries_decoder_generate": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, ctypes.POINTER(GenerateOpts), c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "aries_decoder_detect_language": (c_int, [c_void_p, c_void_p, c_int, ctypes.POINTER(GenerateOpts), c_void_p, c_int,
                                              c_void_p, c_void_p]),
    "aries_decoder_last_stats": (c_int, [c_void_p, c_float_p, c_int]),
    "aries_vad_num_windows": (c_int64, [c_int64]),
    "aries_vad_speech_timestamps": (c_int, [c_void_p, c_int64, c_int64, ctypes.POINTER(VadOpts), c_void_p, c_void_p, c_int,
                                            ctypes.POINTER(c_int)]),
    "aries_vad_energy_probs": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "aries_collect_chunks": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int64,
                                     ctypes.POINTER(c_int64), c_void_p]),
    "aries_test_decoder_generate": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, ctypes.POINTER(GenerateOpts),
                                            c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "aries_test_skinny_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_void_p]),
    "aries_test_skinny_gemm_folded": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "aries_test_skinny_gemm_ln": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "aries_test_decode_attention": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p,
                                            c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                            c_void_p]),
    "aries_test_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "aries_test_gemm_ln": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "aries_test_gemm_stats_parts": (c_int, [c_int]),
    "aries_test_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "aries_test_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "aries_test_attention_trace": (c_int, [c_void_p, c_void_p, c_size_t]),
}

KERNEL_CLASSES = ["logmel_tiles", "logmel_clamp", "mel_transpose", "conv1_gemm", "conv2_gemm", "layernorm", "qkv_gemm",
                  "attention", "oproj_gemm", "fc1_gemm", "fc2_gemm"]

_lib = None
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """Load the library once; raise (never fall back) if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the CUDA library is not built and whisper_aries_b200 has no CPU "
                    "fallback. Run: python -c \"import __graft_entry__ as g; g.build()\"")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def last_error() -> str:
    return load().aries_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map the C ABI's error convention onto upstream's exceptions (ValueError for bad shapes, RuntimeError else)."""
    if rc == ARIES_OK:
        return
    msg = last_error()
    if rc == ARIES_EINVAL:
        raise ValueError(msg)
    if rc == ARIES_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(f"aries_b200 error {rc}: {msg}")


class Context:
    """One ``aries_ctx`` per GPU, shared by every handle created on that GPU."""

    _by_device: dict[int, "Context"] = {}
    _guard = threading.Lock()

    def __init__(self, device_index: int):
        self.lib = load()
        self.device_index = int(device_index)
        h = c_void_p()
        check(self.lib.aries_init(self.device_index, ctypes.byref(h)))
        self.handle = h
        self.sm_count = self.lib.aries_sm_count(h)

    @classmethod
    def get(cls, device_index: int = 0) -> "Context":
        with cls._guard:
            ctx = cls._by_device.get(int(device_index))
            if ctx is None:
                ctx = cls(device_index)
                cls._by_device[int(device_index)] = ctx
            return ctx


def device_index_of(device) -> int:
    """'cuda', 'cuda:3', 3, torch.device('cuda', 3) -> 3.  Anything that is not CUDA is an error (no CPU path)."""
    if isinstance(device, int):
        return device
    s = str(device)
    if s == "cuda" or s == "auto":
        return 0
    if s.startswith("cuda:"):
        return int(s.split(":", 1)[1])
    raise ValueError(f"whisper_aries_b200 runs on CUDA devices only (got device={device!r}); there is no CPU fallback")
