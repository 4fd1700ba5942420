"""Shared helpers for the parity tests (oracle side is CPU only)."""
import numpy as np

from oracle import search_ref, zipformer_ref
from sherpa_vietnamese_asr_b200 import weights


def oracle_recognizer(paths, beam=4, graph=None, dtype=None):
    import torch
    cfgd, tensors = {}, {}
    for part in ("encoder", "decoder", "joiner"):
        c, t = weights.load_container(paths[part])
        cfgd.update(c)
        tensors.update(t)
    cfg = weights.config_from_dict(cfgd)
    id2token = {}
    with open(paths["tokens"], encoding="utf-8") as f:
        for line in f:
            p = line.strip().split()
            if len(p) >= 2:
                id2token[int(p[-1])] = p[0]
    rec = zipformer_ref.make_recognizer(tensors, cfg, id2token=id2token, max_active_paths=beam, context_graph=graph,
                                        dtype=dtype or torch.float32)
    return rec, cfg, tensors


def make_graph(seqs, scores):
    g = search_ref.ContextGraph()
    g.build(seqs, scores)
    return g


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))
