"""CPU oracle for the offline Zipformer RNN-T transcription path.

TEST INFRASTRUCTURE ONLY. Nothing under `oracle/` is part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it, and
there only as the checker (or as the timed CPU stand-in), never inside the CUDA engine's call path.

  fbank_ref.py     Kaldi fbank (kaldi-native-fbank semantics; pinned vs torchaudio + golden vectors)
  zipformer_ref.py Zipformer2 encoder / decoder / joiner in PyTorch-CPU ("parity unpinned": the real
                   ONNX graphs and onnxruntime are unavailable offline)
  search_ref.py    modified_beam_search, greedy, ContextGraph, token statistics, word merge, ROVER
                   (pinned vs the reference's own Python, imported from /root/reference when present,
                   and by tests/golden/* generated from it)
  make_golden.py   regenerates tests/golden/* (needs /root/reference)
"""
