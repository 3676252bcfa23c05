"""whisper_aries_b200 — B200-native (sm_100a) log-mel + Whisper-encoder hot path of Whisper-Aries.

Python host code over a C-ABI CUDA library (include/aries_b200.h).  No CPU fallback: the CUDA library must be built
(``__graft_entry__.build()``) and a B200 must be present for anything but construction-time host logic."""
from . import ct2_model, vad
from .decoder import WhisperDecoder, WhisperGenerationResult
from .encoder import WhisperEncoder, WhisperModel, pcm_s16_to_f32
from .feature_extractor import FeatureExtractor
from .scheduler import (ChunkResult, ChunkScheduler, ChunkWork, Job, chunk_windows, gpu_transcribe_worker, gpu_worker,
                        partition_windows, plan_reference_chunks)
from .synthetic import DEC_SHAPES, SHAPES, DecoderShape, EncoderShape, WhisperTokens

__all__ = ["FeatureExtractor", "WhisperEncoder", "WhisperModel", "ChunkScheduler", "ChunkWork", "ChunkResult", "Job",
           "partition_windows", "plan_reference_chunks", "chunk_windows", "gpu_worker", "gpu_transcribe_worker", "pcm_s16_to_f32", "SHAPES", "EncoderShape", "ct2_model", "vad", "WhisperDecoder", "WhisperGenerationResult", "DEC_SHAPES",
           "DecoderShape", "WhisperTokens"]
