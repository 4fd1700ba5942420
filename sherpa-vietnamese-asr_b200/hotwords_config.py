"""Recognizer-level hotword configuration: from the application's `hotword.txt` to the keyword
arguments `OfflineRecognizer.from_transducer` takes.

Host-side restatement (own code, same behaviour) of /root/reference core/config.py:
  ensure_bpe_vocab       :283-326   bpe.vocab (`piece<TAB>score` per line) written next to bpe.model when it is missing
  prepare_hotwords_file  :329-382   blank lines and `#` comments dropped, `PHRASE :score` kept, phrases upper-cased (the BPE
                                    vocabulary is upper-case), result in a fresh temp file
  get_hotwords_config    :385-414   {} when there are no hotwords, else hotwords_file / hotwords_score 1.5 (+ modeling_unit
                                    "bpe" and bpe_vocab when a vocabulary exists)
Parity: tests/test_host_logic.py runs these beside the reference's functions on the same files.
"""
from __future__ import annotations

import os
import tempfile
from typing import Dict, List

DEFAULT_HOTWORDS_SCORE = 1.5


def ensure_bpe_vocab(model_path: str) -> str:
    model = os.path.join(model_path, "bpe.model")
    vocab = os.path.join(model_path, "bpe.vocab")
    if os.path.exists(vocab):
        return vocab
    if not os.path.exists(model):
        return ""
    try:
        import sentencepiece as spm
        sp = spm.SentencePieceProcessor(model_file=model)
        with open(vocab, "w", encoding="utf-8") as f:
            f.writelines(f"{sp.IdToPiece(i)}\t{sp.GetScore(i)}\n" for i in range(sp.GetPieceSize()))
        return vocab
    except Exception:  # noqa: BLE001   missing sentencepiece or an unreadable model: no BPE hotwords, as in the reference
        return ""


def clean_hotword_lines(text: str) -> List[str]:
    out = []
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        score = ""
        if ":" in line:
            head, tail = line.rsplit(":", 1)
            try:
                float(tail.strip())
                line, score = head.strip(), " :" + tail.strip()
            except ValueError:
                pass
        out.append(line.upper() + score)
    return out


def prepare_hotwords_file(hotwords_path: str, base_dir: str) -> str:
    path = hotwords_path or os.path.join(base_dir, "hotword.txt")
    if not os.path.exists(path):
        return ""
    try:
        with open(path, "r", encoding="utf-8") as f:
            lines = clean_hotword_lines(f.read())
        if not lines:
            return ""
        fd, cleaned = tempfile.mkstemp(suffix=".txt", prefix="asr_hotword_")
        with os.fdopen(fd, "w", encoding="utf-8") as f:
            f.write("\n".join(lines))
        return cleaned
    except Exception:  # noqa: BLE001
        return ""


def get_hotwords_config(model_path: str, base_dir: str) -> Dict[str, object]:
    hotwords_file = prepare_hotwords_file("", base_dir)
    if not hotwords_file:
        return {}
    config: Dict[str, object] = {"hotwords_file": hotwords_file, "hotwords_score": DEFAULT_HOTWORDS_SCORE}
    vocab = ensure_bpe_vocab(model_path)
    if vocab:
        config["modeling_unit"] = "bpe"
        config["bpe_vocab"] = vocab
    return config
