"""TEST INFRASTRUCTURE ONLY - CPU restatement of the voice-activity network the reference runs one window at a time through
onnxruntime (/root/reference core/vad_utils.py:62-118: `session.run(None, {'input': [1, 64+512], 'state': [2,1,128], 'sr'})`
per 512-sample window, 64-sample context and LSTM state carried between calls, state reset per recording).

The model file (`silero_vad.onnx`, Silero VAD v5, 16 kHz branch) is a third-party artefact that is not in /root/reference and
not available offline; its graph is restated here from the published architecture:
    reflect-pad 64 on the right -> STFT as a strided Conv1d (258 filters = real / imaginary parts of 129 bins, kernel 256,
    hop 128, Hann window) -> magnitude [129, 4] -> Conv1d(129->128, k3, p1) ReLU -> Conv1d(128->64, k3, s2, p1) ReLU ->
    Conv1d(64->64, k3, s2, p1) ReLU -> Conv1d(64->128, k3, p1) ReLU -> [128, 1] -> LSTMCell(128, 128) ->
    ReLU -> Conv1d(128->1, k1) -> sigmoid -> mean over time (one frame).
PARITY UNPINNED against the real graph (no weights, no onnxruntime here): pinned structurally (shapes, state carry, window /
context layout of the reference's loop) and used as the oracle of csrc/vad.cu on the same seeded random weights.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

WINDOW, CONTEXT = 512, 64


def stft_basis() -> np.ndarray:
    """[258, 256]: rows 0..128 = cos(2 pi k n / 256) * hann[n], rows 129..257 = -sin(...) * hann[n] (periodic Hann)."""
    n = np.arange(256, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / 256.0)
    k = np.arange(129, dtype=np.float64)[:, None]
    ang = 2.0 * np.pi * k * n[None, :] / 256.0
    return np.concatenate([np.cos(ang) * hann, -np.sin(ang) * hann], axis=0).astype(np.float32)


def init_weights(seed: int = 5) -> dict:
    """Seeded random weights of the architecture above (tensor names as the `.b200w` container stores them). Scales keep the
    activations O(1) and the recurrence contractive, and make the output depend visibly on the input energy."""
    rng = np.random.default_rng(seed)
    W = {"vad.stft.basis": stft_basis()}

    def conv(name, co, ci, k, gain):
        W[name + ".weight"] = (rng.standard_normal((co, ci, k)) * (gain / np.sqrt(ci * k))).astype(np.float32)
        W[name + ".bias"] = (rng.standard_normal(co) * 0.05).astype(np.float32)

    conv("vad.enc0", 128, 129, 3, 0.6)
    conv("vad.enc1", 64, 128, 3, 1.4)
    conv("vad.enc2", 64, 64, 3, 1.4)
    conv("vad.enc3", 128, 64, 3, 1.4)
    W["vad.lstm.weight_ih"] = (rng.standard_normal((512, 128)) * (1.0 / np.sqrt(128))).astype(np.float32)
    W["vad.lstm.weight_hh"] = (rng.standard_normal((512, 128)) * (0.6 / np.sqrt(128))).astype(np.float32)
    W["vad.lstm.bias_ih"] = (rng.standard_normal(512) * 0.05).astype(np.float32)
    W["vad.lstm.bias_hh"] = (rng.standard_normal(512) * 0.05).astype(np.float32)
    W["vad.out.weight"] = (rng.standard_normal((1, 128, 1)) * 0.8).astype(np.float32)
    W["vad.out.bias"] = np.array([-0.3], dtype=np.float32)
    return W


def frontend(W: dict, rows: torch.Tensor) -> torch.Tensor:
    """rows [n, 576] (context + window) -> [n, 128]: everything before the recurrence (window-parallel)."""
    t = lambda k: torch.from_numpy(np.ascontiguousarray(W[k])).to(rows.dtype)
    x = F.pad(rows.unsqueeze(1), (0, 64), mode="reflect")                       # [n, 1, 640]
    s = F.conv1d(x, t("vad.stft.basis").unsqueeze(1), stride=128)               # [n, 258, 4]
    mag = torch.sqrt(s[:, :129] ** 2 + s[:, 129:] ** 2)                         # [n, 129, 4]
    h = F.relu(F.conv1d(mag, t("vad.enc0.weight"), t("vad.enc0.bias"), padding=1))
    h = F.relu(F.conv1d(h, t("vad.enc1.weight"), t("vad.enc1.bias"), stride=2, padding=1))
    h = F.relu(F.conv1d(h, t("vad.enc2.weight"), t("vad.enc2.bias"), stride=2, padding=1))
    h = F.relu(F.conv1d(h, t("vad.enc3.weight"), t("vad.enc3.bias"), padding=1))
    return h[:, :, 0]


def probs(W: dict, rows: np.ndarray, dtype=torch.float32) -> np.ndarray:
    """rows [n, 576] of ONE recording in order (core/vad_utils.py:97-106 layout) -> speech probabilities [n]; the LSTM state
    starts at zero and is carried from window to window as the reference carries `state`."""
    t = lambda k: torch.from_numpy(np.ascontiguousarray(W[k])).to(dtype)
    with torch.no_grad():
        x = frontend(W, torch.from_numpy(np.ascontiguousarray(rows)).to(dtype))
        gx = x @ t("vad.lstm.weight_ih").T + t("vad.lstm.bias_ih") + t("vad.lstm.bias_hh")      # [n, 512] (i, f, g, o)
        whh = t("vad.lstm.weight_hh")
        wo, bo = t("vad.out.weight").reshape(128), t("vad.out.bias")
        h = torch.zeros(128, dtype=dtype)
        c = torch.zeros(128, dtype=dtype)
        out = np.zeros(rows.shape[0], dtype=np.float64)
        for i in range(rows.shape[0]):
            g = gx[i] + whh @ h
            ig, fg, gg, og = torch.sigmoid(g[:128]), torch.sigmoid(g[128:256]), torch.tanh(g[256:384]), torch.sigmoid(g[384:])
            c = fg * c + ig * gg
            h = og * torch.tanh(c)
            out[i] = float(torch.sigmoid(torch.relu(h) @ wo + bo))
    return out.astype(np.float32)


def prob_fn(W: dict):
    """The `vad.get_vad_segments(prob_fn=...)` seam over this oracle."""
    return lambda rows: probs(W, rows)
