"""CPU checks of the voice-activity network's oracle and weights (the GPU kernels are pinned to this oracle in
tests/test_gpu_api.py): tensor inventory, the window / context / state-carry contract of the reference's loop
(core/vad_utils.py:80-106), and the `prob_fn` seam."""
import numpy as np

from oracle import silero_ref
from sherpa_vietnamese_asr_b200 import vad, weights


def test_product_and_oracle_weights_are_the_same_tensors():
    a, b = weights.init_vad_weights(5), silero_ref.init_weights(5)
    assert sorted(a) == sorted(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert a["vad.stft.basis"].shape == (258, 256) and a["vad.lstm.weight_hh"].shape == (512, 128)
    # the basis is a Hann-windowed DFT: bin 0 = the window itself, imaginary part of bin 0 = 0
    np.testing.assert_allclose(a["vad.stft.basis"][0], 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(256) / 256), atol=1e-6)
    assert np.abs(a["vad.stft.basis"][129]).max() == 0.0


def test_container_roundtrip(tmp_path):
    W = weights.init_vad_weights(5)
    path = weights.save_vad(str(tmp_path / "v.b200w"), W)
    cfg, T = weights.load_container(path)
    assert sorted(T) == sorted(W)
    for k in W:
        np.testing.assert_array_equal(T[k], W[k])


def test_oracle_follows_the_reference_loop_contract():
    """State and context are carried from window to window and reset per recording: the second half of a recording decoded
    alone differs from its values inside the whole recording, and the first windows agree."""
    W = silero_ref.init_weights(5)
    rng = np.random.default_rng(0)
    audio = (rng.standard_normal(512 * 40) * 0.1).astype(np.float32)
    audio[512 * 10:512 * 20] *= 0.01
    full = silero_ref.probs(W, vad.window_matrix(audio))
    assert full.shape == (40,) and np.all((full > 0) & (full < 1))
    head = silero_ref.probs(W, vad.window_matrix(audio[:512 * 12]))
    np.testing.assert_array_equal(full[:12], head)
    tail = silero_ref.probs(W, vad.window_matrix(audio[512 * 20:]))
    assert np.abs(tail - full[20:]).max() > 1e-4
    # row layout: 64 samples of context (zeros for the first window) + the window
    rows = vad.window_matrix(audio)
    assert rows.shape == (40, 576) and not rows[0, :64].any()
    np.testing.assert_array_equal(rows[5, :64], audio[512 * 5 - 64:512 * 5])


def test_prob_fn_seam_drives_get_vad_segments():
    W = silero_ref.init_weights(5)
    rng = np.random.default_rng(1)
    audio = (rng.standard_normal(16000 * 6) * 0.05).astype(np.float32)
    segs, probs = vad.get_vad_segments(audio, silero_ref.prob_fn(W))
    assert probs is not None and len(probs) == len(audio) // 512
    assert all(0 <= s < e <= len(audio) for s, e in segs)
