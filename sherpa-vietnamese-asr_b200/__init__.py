"""B200-native offline Zipformer RNN-T transcription engine (drop-in for the reference's recognizer
surface). Host code is Python over a C-ABI CUDA library (`csrc/` -> `libb200asr.so`, ctypes)."""
__version__ = "0.1.0"
