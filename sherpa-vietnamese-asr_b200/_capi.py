"""ctypes binding of libb200asr.so (include/b200asr.h). No torch types cross this boundary.

The library is built in-tree by `build.py`; if it is missing or cannot be loaded this module raises —
there is deliberately no CPU fallback (the product path is the CUDA path).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200asr.so")


class FeatureConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("feature_dim", C.c_int32)]


class TransducerModelConfig(C.Structure):
    _fields_ = [("encoder", C.c_char_p), ("decoder", C.c_char_p), ("joiner", C.c_char_p)]


class ModelConfig(C.Structure):
    _fields_ = [("transducer", TransducerModelConfig), ("tokens", C.c_char_p), ("num_threads", C.c_int32),
                ("debug", C.c_int32), ("provider", C.c_char_p), ("model_type", C.c_char_p),
                ("modeling_unit", C.c_char_p), ("bpe_vocab", C.c_char_p)]


class RecognizerConfig(C.Structure):
    _fields_ = [("feat_config", FeatureConfig), ("model_config", ModelConfig), ("decoding_method", C.c_char_p),
                ("max_active_paths", C.c_int32), ("hotwords_file", C.c_char_p), ("hotwords_score", C.c_float),
                ("blank_penalty", C.c_float), ("device_id", C.c_int32), ("precision", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("text", C.c_char_p), ("json", C.c_char_p), ("tokens", C.POINTER(C.c_char_p)),
                ("token_ids", C.POINTER(C.c_int32)), ("timestamps", C.POINTER(C.c_float)),
                ("frames", C.POINTER(C.c_int32)), ("ys_log_probs", C.POINTER(C.c_float)),
                ("tsallis", C.POINTER(C.c_float)), ("margin", C.POINTER(C.c_float)),
                ("entropy", C.POINTER(C.c_float)), ("top1", C.POINTER(C.c_float)), ("count", C.c_int32),
                ("num_frames", C.c_int32), ("duration", C.c_float)]


# name -> (restype, argtypes); every symbol include/b200asr.h declares
_P = C.c_void_p
_F = C.POINTER(C.c_float)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)
SYMBOLS = {
    "B200AsrCreateOfflineRecognizer": (_P, [C.POINTER(RecognizerConfig)]),
    "B200AsrDestroyOfflineRecognizer": (None, [_P]),
    "B200AsrOfflineRecognizerSetConfig": (C.c_int32, [_P, C.POINTER(RecognizerConfig)]),
    "B200AsrSetHotwordsTokenIds": (C.c_int32, [_P, _I32, _I32, _F, C.c_int32]),
    "B200AsrCreateOfflineStream": (_P, [_P]),
    "B200AsrDestroyOfflineStream": (None, [_P]),
    "B200AsrAcceptWaveformOffline": (C.c_int32, [_P, C.c_int32, _F, C.c_int32]),
    "B200AsrAcceptFeaturesOffline": (C.c_int32, [_P, _F, C.c_int32, C.c_int32, C.c_int64]),
    "B200AsrAcceptWaveformsOffline": (C.c_int32, [C.POINTER(_P), C.c_int32, C.POINTER(_P), _I32, C.c_int32]),
    "B200AsrCreateOfflineStreamWithHotwords": (_P, [_P, C.c_char_p]),
    "B200AsrDecodeOfflineStream": (C.c_int32, [_P, _P]),
    "B200AsrDecodeMultipleOfflineStreams": (C.c_int32, [_P, C.POINTER(_P), C.c_int32]),
    "B200AsrGetOfflineStreamResult": (C.POINTER(Result), [_P]),
    "B200AsrDestroyOfflineRecognizerResult": (None, [C.POINTER(Result)]),
    "B200AsrGetOfflineStreamResultAsJson": (_P, [_P]),
    "B200AsrDestroyOfflineStreamResultJson": (None, [_P]),
    "B200AsrGetLastError": (C.c_char_p, []),
    "B200AsrVersion": (C.c_char_p, []),
    "B200AsrVocabSize": (C.c_int32, [_P]),
    "B200AsrEncoderOutDim": (C.c_int32, [_P]),
    "B200AsrFbank": (C.c_int32, [_P, _F, C.c_int32, _F]),
    "B200AsrFbankBatch": (C.c_int32, [_P, _F, _I64, C.c_int32, _F, _I64]),
    "B200AsrSilentFrames": (C.c_int32, [_F, C.c_int64, C.c_int32, C.c_float, C.POINTER(C.c_uint8), C.c_int32]),
    "B200AsrPreprocessAudio": (C.c_int32, [_F, C.c_int64, _I64, _I64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F, C.c_int32]),
    "B200AsrVadCreate": (_P, [C.c_char_p, C.c_int32]),
    "B200AsrVadDestroy": (None, [_P]),
    "B200AsrVadProbs": (C.c_int32, [_P, _F, C.c_int64, _F]),
    "B200AsrVadProbsBatch": (C.c_int32, [_P, _F, _I64, C.c_int32, _F, _I64]),
    "B200AsrVadLastTimings": (C.c_int32, [_P, _F, _F]),
    "B200AsrEncoder": (C.c_int32, [_P, _F, _I32, C.c_int32, _F, _I32]),
    "B200AsrEncoderTap": (C.c_int32, [_P, C.c_char_p, _F, _I32]),
    "B200AsrDecoder": (C.c_int32, [_P, _I64, C.c_int32, _F]),
    "B200AsrJoiner": (C.c_int32, [_P, _F, _F, C.c_int32, _F]),
    "B200AsrDecoderJoinerInput": (C.c_int32, [_P, _I64, _F, C.c_int32, _F, _F]),
    "B200AsrJoinerRecords": (C.c_int32, [_P, _F, C.c_int32, C.c_int32, _F]),
    "B200AsrBeamSearch": (C.c_int32, [_P, _F, _I32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I32, _I32, _F, _F, _I32]),
    "B200AsrGemm": (C.c_int32, [_P, _F, _F, _F, _F, _F, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F]),
    "B200AsrContextForwardOneStep": (C.c_double, [_P, C.c_int32, C.c_int32, _I32]),
    "B200AsrContextFinalize": (C.c_double, [_P, C.c_int32]),
    "B200AsrContextNumNodes": (C.c_int32, [_P]),
    "B200AsrHotwordGraphCreate": (_P, [_I32, _I32, _F, C.c_int32]),
    "B200AsrHotwordGraphDestroy": (None, [_P]),
    "B200AsrHotwordGraphNumNodes": (C.c_int32, [_P]),
    "B200AsrHotwordGraphStep": (C.c_double, [_P, C.c_int32, C.c_int32, _I32]),
    "B200AsrHotwordGraphFinalize": (C.c_double, [_P, C.c_int32]),
    "B200AsrStageBatch": (C.c_int32, [_P, _F, _I64, C.c_int32]),
    "B200AsrRunStagedBatch": (C.c_int32, [_P, C.c_int32, _I32]),
    "B200AsrRunStagedBatchChained": (C.c_int32, [_P, C.c_int32, C.c_int32, _I32, _F]),
    "B200AsrReleaseBatch": (C.c_int32, [_P, C.c_int32]),
    "B200AsrLastPassTokens": (C.c_int32, [_P, C.c_int32, _I32, _I32, C.c_int32]),
    "B200AsrLastPipelineStats": (C.c_int32, [_P, _I32, _F, _F, _I64]),
    "B200AsrLastPipelineTimeline": (C.c_int32, [_P, _F, C.c_int32]),
    "B200AsrLastTimings": (C.c_int32, [_P, _F, _I64]),
    "B200AsrLastGemmStats": (C.c_int32, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), _I64]),
    "B200AsrLastGemmBytes": (C.c_double, [_P]),
    "B200AsrSetProfiling": (C.c_int32, [_P, C.c_int32]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads libb200asr.so (once). Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                               "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def last_error() -> str:
    return (lib().B200AsrGetLastError() or b"").decode("utf-8", "replace")


def fptr(a):
    return a.ctypes.data_as(_F)


def i32ptr(a):
    return a.ctypes.data_as(_I32)


def i64ptr(a):
    return a.ctypes.data_as(_I64)
