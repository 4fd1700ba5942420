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


def rel_l2(a, b):
    """||a - b|| / ||b|| — the output-difference metric of the reference's own CPU-vs-GPU gate
    (/root/reference core/calibration.py:1057-1090 `_output_diff`)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    if a.size == 0:
        return 0.0
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


def row_err(a, b):
    """Largest per-row error relative to that row's own scale: a localised fault cannot hide behind the tensor's global maximum."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    a = a.reshape(-1, a.shape[-1])
    b = b.reshape(-1, b.shape[-1])
    return float((np.abs(a - b).max(axis=1) / (np.abs(b).max(axis=1) + 1e-6)).max())
