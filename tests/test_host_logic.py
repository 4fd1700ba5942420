"""CPU suite for the product's host side: C-ABI surface, hotword parsing/tokenising, token->word merge and
ROVER (against golden vectors made by the reference), workload generators, sharding."""
import copy
import json
import os
import re

import numpy as np
import pytest

from oracle import search_ref as sr, zipformer_ref as zr
from sherpa_vietnamese_asr_b200 import _capi, asr_engine, recognizer, synth, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200asr.h")).read()
    declared = set(re.findall(r"B200ASR_API[^;]*?\b(B200Asr\w+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    lib = _capi.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.B200AsrVersion()


def test_no_cpu_fallback(tmp_path):
    """Without a CUDA device the recognizer must fail loudly, never fall back."""
    import ctypes
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    paths = weights.write_model_dir(str(tmp_path), weights.zipformer_tiny(), 3)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        recognizer.OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"],
                                                     joiner=paths["joiner"], tokens=paths["tokens"])
    cfg = _capi.RecognizerConfig()
    assert not _capi.lib().B200AsrCreateOfflineRecognizer(ctypes.byref(cfg))
    assert _capi.last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sherpa-vietnamese-asr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_parse_hotwords_text_matches_reference_rules():
    txt = "# comment\n\nHà Nội :2.5\nban tổ chức\nA:B\nfoo :x\n  spaced phrase  : 1.0 \n"
    got = recognizer.parse_hotwords_text(txt, 1.5)
    assert got == [("HÀ NỘI", 2.5), ("BAN TỔ CHỨC", 1.5), ("A:B", 1.5), ("FOO :X", 1.5), ("SPACED PHRASE", 1.0)]


def test_table_tokenizer_roundtrip():
    cfg = weights.zipformer_tiny()
    toks = weights.make_tokens(cfg)
    tt = recognizer._TableTokenizer(toks)
    word_starts = [t for t in toks[3:] if t.startswith("▁")]
    phrase = " ".join(t[1:] for t in word_starts[:3])
    ids = tt.encode(phrase)
    assert ids and "".join(toks[i] for i in ids).replace("▁", " ").strip() == phrase
    assert tt.encode("@@@") == []


class _Res:
    pass


def _result_from_oracle(cfg, W, id2token, enc):
    rec = zr.make_recognizer(W, cfg, id2token=id2token, max_active_paths=4)
    toks, frames, lps, T, emit = sr.modified_beam_search(rec, None, 4, enc_out=enc)
    r = _Res()
    r.token_ids, r.frames, r.ys_log_probs, r.num_frames = toks, frames, lps, T
    st = [sr.token_entropy(e, cfg.vocab_size, rounded=False) for e in emit]
    r.tsallis = [s["tsallis_norm"] for s in st]
    r.margin = [s["margin"] for s in st]
    r.entropy = [s["entropy_norm"] for s in st]
    r.top1 = [s["top1_prob"] for s in st]
    return r


def test_words_from_result_matches_reference_golden():
    cfg = weights.zipformer_tiny()
    W = weights.init_weights(cfg, 3)
    id2token = {i: t for i, t in enumerate(weights.make_tokens(cfg))}
    encs = np.load(os.path.join(GOLD, "words_enc.npz"))
    for c in json.load(open(os.path.join(GOLD, "words.json"), encoding="utf-8")):
        res = _result_from_oracle(cfg, W, id2token, encs[f"enc_{c['id']}"])
        words = asr_engine.words_from_result(res, id2token, c["n_samples"], c["time_offset"])
        assert len(words) == len(c["words"])
        for got, want in zip(words, c["words"]):
            assert set(got) == set(want)
            for k, v in want.items():
                if isinstance(v, float):
                    assert got[k] == pytest.approx(v, abs=1e-9), k
                else:
                    assert got[k] == v, k


def test_product_rover_matches_reference_golden():
    for c in json.load(open(os.path.join(GOLD, "rover.json"), encoding="utf-8")):
        merged, dis = asr_engine.rover_merge_words(copy.deepcopy(c["a"]), copy.deepcopy(c["b"]), c["hotword_phrases"])
        assert json.loads(json.dumps(merged)) == c["merged"]
        assert sorted(dis) == c["disagree"]


def test_synthetic_workloads_are_seeded_and_shaped():
    a, b = synth.c1_clip(), synth.c1_clip()
    assert a.shape == (160000,) and a.dtype == np.float32 and np.array_equal(a, b)
    assert 0.55 <= np.abs(a).max() <= 1.0
    d = synth.c2_durations()
    assert d.shape == (256,) and d.min() >= 1.0 and d.max() <= 30.0 and 2400 < d.sum() < 3400
    seqs, scores = synth.random_hotwords(500, 2000, 500)
    assert len(seqs) == 500 and all(2 <= len(s) <= 8 for s in seqs) and set(scores) <= {1.5, 2.0, 2.5}
    assert synth.speech_like(1, 0).shape == (1,)


def test_create_recognizer_discovery_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        asr_engine.create_recognizer(str(tmp_path))


def test_result_marshalling_from_the_c_struct():
    """OfflineStream.result copies the library-owned result arrays into Python lists: same values and Python types as
    element-wise ctypes reads (float32 values widened to Python floats, int32 to ints), empty results included."""
    import ctypes as C
    from sherpa_vietnamese_asr_b200.recognizer import result_from_struct
    n = 7
    rng = np.random.default_rng(0)
    ids = (C.c_int32 * n)(*[int(x) for x in rng.integers(0, 2000, n)])
    frames = (C.c_int32 * n)(*range(3, 3 + n))
    f = lambda: (C.c_float * n)(*[float(x) for x in rng.normal(0, 1, n).astype(np.float32)])
    ts, lps, tsl, mg, en, t1 = f(), f(), f(), f(), f(), f()
    toks = (C.c_char_p * n)(*[("▁t%d" % i).encode("utf-8") for i in range(n)])
    r = _capi.Result(text="xin chào".encode("utf-8"), json=b'{"text": "xin"}', tokens=toks, token_ids=ids, timestamps=ts, frames=frames,
                     ys_log_probs=lps, tsallis=tsl, margin=mg, entropy=en, top1=t1, count=n, num_frames=248, duration=10.0)
    res = result_from_struct(r)
    assert res.text == "xin chào" and res.tokens == ["▁t%d" % i for i in range(n)] and str(res) == '{"text": "xin"}'
    assert res.token_ids == [ids[i] for i in range(n)] and all(type(x) is int for x in res.token_ids)
    assert res.frames == list(range(3, 3 + n)) and res.num_frames == 248 and res.duration == 10.0
    for got, src in ((res.timestamps, ts), (res.ys_log_probs, lps), (res.tsallis, tsl), (res.margin, mg), (res.entropy, en), (res.top1, t1)):
        assert got == [src[i] for i in range(n)] and all(type(x) is float for x in got)
    empty = result_from_struct(_capi.Result(count=0))
    assert empty.token_ids == [] and empty.text == "" and empty.tokens == [] and empty.timestamps == []


# ----------------------------------------------------------------------------- hotword configuration (core/config.py:283-414)
HOTWORD_TXT = """# tên riêng
Hồ Chí Minh :2.5
  thành phố hà nội
trí tuệ nhân tạo:2.0
tỷ lệ 3:1

#bỏ qua
chat gpt : abc
"""


def _tiny_sentencepiece(model_dir):
    spm = pytest.importorskip("sentencepiece")
    corpus = os.path.join(model_dir, "corpus.txt")
    with open(corpus, "w", encoding="utf-8") as f:
        for i in range(200):
            f.write("XIN CHÀO CÁC BẠN HÔM NAY CHÚNG TA NÓI VỀ THÀNH PHỐ HỒ CHÍ MINH VÀ HÀ NỘI TRÍ TUỆ NHÂN TẠO %d\n" % i)
    spm.SentencePieceTrainer.train(input=corpus, model_prefix=os.path.join(model_dir, "bpe"), vocab_size=60, model_type="bpe",
                                   minloglevel=2)
    os.remove(os.path.join(model_dir, "bpe.vocab"))      # the trainer writes one too; ensure_bpe_vocab has to create it


def test_hotword_config_matches_reference(tmp_path):
    from sherpa_vietnamese_asr_b200 import hotwords_config as hw
    base = tmp_path / "app"
    base.mkdir()
    (base / "hotword.txt").write_text(HOTWORD_TXT, encoding="utf-8")
    want_lines = ["HỒ CHÍ MINH :2.5", "THÀNH PHỐ HÀ NỘI", "TRÍ TUỆ NHÂN TẠO :2.0", "TỶ LỆ 3 :1", "CHAT GPT : ABC"]
    assert hw.clean_hotword_lines(HOTWORD_TXT) == want_lines
    cleaned = hw.prepare_hotwords_file("", str(base))
    assert os.path.basename(cleaned).startswith("asr_hotword_") and open(cleaned, encoding="utf-8").read() == "\n".join(want_lines)
    # what the cleaned file means to the recognizer: phrase + score pairs
    assert recognizer.parse_hotwords_text(open(cleaned, encoding="utf-8").read()) == [
        ("HỒ CHÍ MINH", 2.5), ("THÀNH PHỐ HÀ NỘI", 1.5), ("TRÍ TUỆ NHÂN TẠO", 2.0), ("TỶ LỆ 3", 1.0), ("CHAT GPT : ABC", 1.5)]
    assert hw.prepare_hotwords_file("", str(tmp_path / "nowhere")) == ""
    (base / "empty.txt").write_text("# only comments\n\n", encoding="utf-8")
    assert hw.prepare_hotwords_file(str(base / "empty.txt"), str(base)) == ""

    models = {}
    for tag in ("ours", "ref"):
        d = tmp_path / f"model_{tag}"
        d.mkdir()
        _tiny_sentencepiece(str(d))
        models[tag] = str(d)
    assert hw.ensure_bpe_vocab(str(tmp_path / "nowhere")) == ""
    cfg = hw.get_hotwords_config(models["ours"], str(base))
    assert set(cfg) == {"hotwords_file", "hotwords_score", "modeling_unit", "bpe_vocab"}
    assert cfg["hotwords_score"] == 1.5 and cfg["modeling_unit"] == "bpe" and cfg["bpe_vocab"] == os.path.join(models["ours"], "bpe.vocab")
    vocab_lines = open(cfg["bpe_vocab"], encoding="utf-8").read().splitlines()
    assert len(vocab_lines) == 60 and all(len(l.split("\t")) == 2 for l in vocab_lines)
    assert hw.get_hotwords_config(models["ours"], str(tmp_path / "nowhere")) == {}
    no_model = tmp_path / "model_none"
    no_model.mkdir()
    assert set(hw.get_hotwords_config(str(no_model), str(base))) == {"hotwords_file", "hotwords_score"}

    if not os.path.isdir("/root/reference/core"):
        return
    import io, sys
    from contextlib import redirect_stdout
    sys.dont_write_bytecode = True
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    with redirect_stdout(io.StringIO()):
        import core.config as rc
        ref_clean = rc.prepare_hotwords_file("", str(base))
        ref_cfg = rc.get_hotwords_config(models["ref"], str(base))
        assert rc.prepare_hotwords_file(str(base / "empty.txt"), str(base)) == ""
    assert open(ref_clean, encoding="utf-8").read() == open(cleaned, encoding="utf-8").read()
    assert set(ref_cfg) == set(cfg) and ref_cfg["hotwords_score"] == cfg["hotwords_score"] and ref_cfg["modeling_unit"] == cfg["modeling_unit"]
    assert open(ref_cfg["bpe_vocab"], encoding="utf-8").read() == open(cfg["bpe_vocab"], encoding="utf-8").read()
    assert open(ref_cfg["hotwords_file"], encoding="utf-8").read() == open(cfg["hotwords_file"], encoding="utf-8").read()
