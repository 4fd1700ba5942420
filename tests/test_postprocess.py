"""CPU suite for the post-recognizer host steps (suspect flags from the search statistics,
gap features, filler removal, asr_json): our code against golden vectors written by the reference's own functions
(oracle/make_golden.py -> tests/golden/postprocess.json) and against those functions imported live when /root/reference exists."""
import copy
import io
import json
import os
import sys
from contextlib import redirect_stdout

import pytest

from oracle import post_cases as pc
from sherpa_vietnamese_asr_b200 import postprocess as pp

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"
have_ref = os.path.isdir(os.path.join(REF, "core"))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "postprocess.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def ref():
    if not have_ref:
        pytest.skip("/root/reference not present (GPU box)")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with redirect_stdout(io.StringIO()):
        import core.asr_engine as ae
        import core.asr_json as aj
        import core.vad_utils as vu
    return ae, aj, vu


def _summary(words):
    return [[w.get("_suspect_level"), w.get("gap_after_ms"), w.get("gap_before_ms")] for w in words]


# ----------------------------------------------------------------------------- golden
def test_suspect_detect_and_fillers_golden(gold):
    seen_gap = seen_flag = 0
    for (words, audio, disagree, vad), want in zip(pc.suspect_cases(), gold["suspect"]):
        got = pp.suspect_detect(copy.deepcopy(words), audio, disagree, vad)
        assert _summary(got) == want["flags"]
        assert [w["start"] for w in pp.remove_filler_words(got)] == want["kept"]
        seen_gap += sum(1 for f in want["flags"] if f[1])
        seen_flag += sum(1 for f in want["flags"] if f[0])
    assert seen_gap > 20 and seen_flag > 100


def test_gap_features_golden(gold):
    for seg, want in zip(pc.gap_segments(), gold["gap"]):
        assert pp.count_energy_peaks(seg) == want["peaks"]
        assert list(pp.compute_gap_features(seg)) == want["features"]
    assert max(len(w["peaks"]) for w in gold["gap"]) >= 3


def test_disagree_indices_golden(gold):
    for (main, other), want in zip(pc.disagree_cases(), gold["disagree"]):
        assert sorted(pp.compute_disagree_indices(main, other)) == want


def test_asr_json_golden(gold):
    for c, want in zip(pc.segment_cases(), gold["serialize"]):
        data = pp.serialize_segments(copy.deepcopy(c["segments"]), c["speaker_name_mapping"], c["speaker_colors"], c["model_name"],
                                     c["model_type"], c["duration_sec"], c["timing"], c["overlap_segments"])
        assert isinstance(data.pop("created_at"), str)
        data = json.loads(json.dumps(data))                      # what goes over the wire
        assert data == want["json"]
        assert list(data) == list(want["json"])                  # same key order in the file
        back = pp.deserialize_segments(data)
        assert json.loads(json.dumps(list(back))) == want["back"]
    with pytest.raises(ValueError):
        pp.deserialize_segments({"version": 1})


# ----------------------------------------------------------------------------- live against the reference
@pytest.mark.parametrize("seed", [101, 102, 103, 104])
def test_suspect_detect_live(ref, seed):
    ae, _, vu = ref
    stats = ["tsallis+margin", "tsallis", "entropy", "tsallis+margin"][seed % 4]
    words, audio, disagree, vad = pc.suspect_case(seed, 45.0, stats, with_disagree=seed % 2 == 0)
    try:
        vu._last_vad_probs = vad
        with redirect_stdout(io.StringIO()):
            want = ae.suspect_detect(copy.deepcopy(words), audio, disagree_indices=disagree)
            want_kept = ae.remove_filler_words(want)
    finally:
        vu._last_vad_probs = None
    got = pp.suspect_detect(copy.deepcopy(words), audio, disagree, vad)
    assert got == want
    assert pp.remove_filler_words(got) == want_kept


def test_gap_features_live(ref):
    ae = ref[0]
    for seg in pc.gap_segments(77, 60):
        assert pp.count_energy_peaks(seg) == ae.count_energy_peaks(seg)
        assert pp.count_energy_peaks(seg, 16000, 0.5) == ae.count_energy_peaks(seg, 16000, 0.5)
        assert pp.compute_gap_features(seg) == ae._compute_gap_features(seg)


def test_disagree_live(ref):
    ae = ref[0]
    for main, other in pc.disagree_cases(5, 60):
        assert pp.compute_disagree_indices(main, other) == ae.compute_disagree_indices(main, other)


def test_asr_json_live(ref):
    aj = ref[1]
    for c in pc.segment_cases(43):
        args = (c["speaker_name_mapping"], c["speaker_colors"], c["model_name"], c["model_type"], c["duration_sec"], c["timing"],
                c["overlap_segments"])
        got = pp.serialize_segments(copy.deepcopy(c["segments"]), *args)
        want = aj.serialize_segments(copy.deepcopy(c["segments"]), *args)
        got.pop("created_at"), want.pop("created_at")
        assert got == want and json.dumps(got, ensure_ascii=False) == json.dumps(want, ensure_ascii=False)
        assert pp.deserialize_segments(got) == aj.deserialize_segments(want)


# ----------------------------------------------------------------------------- the post-ASR sequence
def test_finish_transcript_sequence():
    words, audio, _, vad = pc.suspect_case(9, 20.0)
    for i in (2, 5):
        words[i]["_disagree"] = True
    ref_words = copy.deepcopy(words)
    out, text = pp.finish_transcript(words, audio, is_rover=True, vad_probs=vad)
    for w in ref_words:
        w.pop("_disagree", None)
    want = pp.remove_filler_words(pp.suspect_detect(ref_words, audio, {2, 5}, vad))
    assert out == want and all("_disagree" not in w for w in out)
    assert text == " ".join(w["text"] for w in want).capitalize() and text[0] == text[0].upper()
    # single-model runs carry no disagreement set
    words2 = pc.suspect_case(9, 20.0)[0]
    words2[2]["_disagree"] = True
    out2, _ = pp.finish_transcript(words2, audio, is_rover=False, vad_probs=vad)
    base = pp.remove_filler_words(pp.suspect_detect(pc.suspect_case(9, 20.0)[0], audio, None, vad))
    assert [w.get("_suspect_level") for w in out2] == [w.get("_suspect_level") for w in base]
    assert pp.finish_transcript([], audio) == ([], "")
