"""Model configs, seeded random-init weights and the `.b200w` weight container.

The reference loads three ONNX graphs per model directory (`encoder-*.onnx`,
`decoder-*.onnx`, `joiner-*.onnx`, /root/reference core/asr_engine.py:912-927).
Those checkpoints are not available offline, so models here are random-init
tensors of the named architecture stored in a
container the C library parses without any third-party dependency.

Container layout (little endian):
    line 1      : "B200ASRW 1 <header_bytes>\n"
    header lines: "config <key> <value>\n"
                  "tensor <name> f32 <ndim> <d0> .. <dn-1> <offset> <nbytes>\n"
                  "end\n"  then zero padding up to <header_bytes> (multiple of 4096)
    payload     : raw tensors, each 256-byte aligned, offsets relative to <header_bytes>
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

MAGIC = "B200ASRW"


@dataclass
class ZipformerConfig:
    name: str
    num_encoder_layers: tuple = (2, 2, 3, 4, 3, 2)
    downsampling_factor: tuple = (1, 2, 4, 8, 4, 2)
    encoder_dim: tuple = (192, 256, 384, 512, 384, 256)
    feedforward_dim: tuple = (512, 768, 1024, 1536, 1024, 768)
    num_heads: tuple = (4, 4, 4, 8, 4, 4)
    cnn_module_kernel: tuple = (31, 31, 15, 15, 15, 31)
    query_head_dim: int = 32
    pos_head_dim: int = 4
    value_head_dim: int = 12
    pos_dim: int = 48
    feature_dim: int = 80
    decoder_dim: int = 512
    joiner_dim: int = 512
    context_size: int = 2
    vocab_size: int = 2000
    blank_id: int = 0
    unk_id: int = 2

    @property
    def output_dim(self) -> int:
        return max(self.encoder_dim)


def zipformer_68m() -> ZipformerConfig:
    """icefall Zipformer2 'medium' = sherpa-onnx-zipformer-vi-2025-04-20."""
    return ZipformerConfig(name="zipformer-68m")


def zipformer_30m() -> ZipformerConfig:
    """icefall Zipformer2 'small' = zipformer-30m-rnnt-6000h."""
    return ZipformerConfig(
        name="zipformer-30m",
        num_encoder_layers=(2, 2, 2, 2, 2, 2),
        encoder_dim=(192, 256, 256, 256, 256, 256),
        feedforward_dim=(512, 768, 768, 768, 768, 768),
    )


def zipformer_tiny() -> ZipformerConfig:
    """A small same-topology model for fast CPU tests (not a reference model)."""
    return ZipformerConfig(
        name="zipformer-tiny",
        num_encoder_layers=(1, 1, 1, 1, 1, 1),
        encoder_dim=(64, 96, 128, 160, 128, 96),
        feedforward_dim=(128, 192, 256, 320, 256, 192),
        num_heads=(4, 4, 4, 8, 4, 4),
        cnn_module_kernel=(31, 31, 15, 15, 15, 31),
        decoder_dim=128, joiner_dim=128, vocab_size=500,
    )


CONFIGS = {"zipformer-68m": zipformer_68m, "zipformer-30m": zipformer_30m,
           "zipformer-tiny": zipformer_tiny}


# --------------------------------------------------------------------------- init
def _linear(rng, out_f, in_f, gain=1.0, bias_std=0.01):
    """Biases are kept small (but non-zero, so bias handling is exercised): constant offsets through 16 residual layers would swamp the
    time-varying part of the signal in an untrained network."""
    w = (rng.standard_normal((out_f, in_f)) * (gain / math.sqrt(in_f))).astype(np.float32)
    b = (rng.standard_normal(out_f) * bias_std).astype(np.float32)
    return w, b


def init_weights(cfg: ZipformerConfig, seed: int) -> dict:
    """Seeded random init (tensor inventory: SURVEY App. B.7).

    Scales are chosen so that activations stay O(1) through the stacks, attention logits have
    unit-ish variance (the 1/sqrt(d) is baked into in_proj as icefall does), and the joiner is
    peaky with a dominant blank so beam search emits sparse, well-separated tokens
    (SURVEY §8d 'Weights').
    """
    rng = np.random.default_rng(seed)
    W: dict[str, np.ndarray] = {}

    def put(name, arr):
        W[name] = np.ascontiguousarray(arr, dtype=np.float32)

    def conv2d(name, co, ci, kh, kw, gain=1.0, zero_sum=False):
        fan = ci * kh * kw
        w = rng.standard_normal((co, ci, kh, kw)) * (gain / math.sqrt(fan))
        if zero_sum:   # DC-blocking kernels: keep the time-varying part of log-mel, drop its offset
            w -= w.mean(axis=(1, 2, 3), keepdims=True)
        put(name + ".weight", w)
        put(name + ".bias", rng.standard_normal(co) * 0.01)

    # ---- encoder_embed (Conv2dSubsampling, App. B.2)
    e = "encoder.embed."
    conv2d(e + "conv0", 8, 1, 3, 3, gain=1.05, zero_sum=True)
    conv2d(e + "conv1", 32, 8, 3, 3, gain=1.6, zero_sum=True)
    conv2d(e + "conv2", 128, 32, 3, 3, gain=1.6, zero_sum=True)
    conv2d(e + "convnext.dw", 128, 1, 7, 7, gain=1.0)
    w, b = _linear(rng, 384, 128, 1.4); put(e + "convnext.pw1.weight", w); put(e + "convnext.pw1.bias", b)
    w, b = _linear(rng, 128, 384, 0.7); put(e + "convnext.pw2.weight", w); put(e + "convnext.pw2.bias", b)
    d0 = cfg.encoder_dim[0]
    out_w = (((cfg.feature_dim - 1) // 2) - 1) // 2
    w, b = _linear(rng, d0, 128 * out_w, 1.0); put(e + "out.weight", w); put(e + "out.bias", b)
    put(e + "out_norm.bias", rng.standard_normal(d0) * 0.05)   # BiasNorm biases stay non-zero
    put(e + "out_norm.log_scale", np.array([0.0]))

    # ---- stacks
    for i, (L, ds, D, F, H, k) in enumerate(zip(cfg.num_encoder_layers, cfg.downsampling_factor,
                                                 cfg.encoder_dim, cfg.feedforward_dim,
                                                 cfg.num_heads, cfg.cnn_module_kernel)):
        s = f"encoder.stack{i}."
        if ds > 1:
            put(s + "downsample.bias", rng.standard_normal(ds) * 0.3)
            put(s + "out_combiner.scale", rng.uniform(0.35, 0.65, D))
        qd, pd, vd = cfg.query_head_dim, cfg.pos_head_dim, cfg.value_head_dim
        for l in range(L):
            p = s + f"layer{l}."
            w, b = _linear(rng, H * (2 * qd + pd), D, 1.0)
            w[: 2 * H * qd] *= qd ** -0.25 * 1.5          # q,k rows: scaled dot product baked in
            b[: 2 * H * qd] *= qd ** -0.25
            w[2 * H * qd:] *= 0.6
            b[2 * H * qd:] = 1.0                           # constant positional query ...
            put(p + "attn_w.in_proj.weight", w); put(p + "attn_w.in_proj.bias", b)
            lp = rng.standard_normal((H * pd, cfg.pos_dim)) * (0.24 / math.sqrt(cfg.pos_dim))
            lp[:, 0:cfg.pos_dim - 1:2] += 0.25            # ... times a cosine comb peaking at offset 0:
            put(p + "attn_w.linear_pos.weight", lp)       # attention prefers nearby frames, as trained models do
            for j, f in enumerate(((F * 3) // 4, F, (F * 5) // 4), start=1):
                w, b = _linear(rng, f, D, 1.2); put(p + f"ff{j}.in.weight", w); put(p + f"ff{j}.in.bias", b)
                w, b = _linear(rng, D, f, 0.5); put(p + f"ff{j}.out.weight", w); put(p + f"ff{j}.out.bias", b)
            h = (3 * D) // 4
            w, b = _linear(rng, 3 * h, D, 1.0); put(p + "nonlin.in.weight", w); put(p + "nonlin.in.bias", b)
            w, b = _linear(rng, D, h, 0.7); put(p + "nonlin.out.weight", w); put(p + "nonlin.out.bias", b)
            for j in (1, 2):
                w, b = _linear(rng, H * vd, D, 1.0); put(p + f"attn{j}.in.weight", w); put(p + f"attn{j}.in.bias", b)
                w, b = _linear(rng, D, H * vd, 0.6); put(p + f"attn{j}.out.weight", w); put(p + f"attn{j}.out.bias", b)
                w, b = _linear(rng, 2 * D, D, 1.0); put(p + f"conv{j}.in.weight", w); put(p + f"conv{j}.in.bias", b)
                put(p + f"conv{j}.dw.weight", rng.standard_normal((D, 1, k)) * (1.2 / math.sqrt(k)))
                put(p + f"conv{j}.dw.bias", rng.standard_normal(D) * 0.01)
                w, b = _linear(rng, D, D, 0.6); put(p + f"conv{j}.out.weight", w); put(p + f"conv{j}.out.bias", b)
            put(p + "norm.bias", rng.standard_normal(D) * 0.05)
            put(p + "norm.log_scale", np.array([rng.uniform(-0.1, 0.1)]))
            put(p + "bypass.scale", rng.uniform(0.4, 0.9, D))
            put(p + "bypass_mid.scale", rng.uniform(0.4, 0.9, D))
    put("encoder.downsample_output.bias", rng.standard_normal(2) * 0.3)
    w, b = _linear(rng, cfg.joiner_dim, cfg.output_dim, 2.6)
    put("encoder.encoder_proj.weight", w); put("encoder.encoder_proj.bias", b)

    # ---- decoder (stateless, App. B.4)
    dd = cfg.decoder_dim
    put("decoder.embedding.weight", rng.standard_normal((cfg.vocab_size, dd)))
    put("decoder.conv.weight", rng.standard_normal((dd, 4, cfg.context_size)) * (1.0 / math.sqrt(4 * cfg.context_size)))
    w, b = _linear(rng, cfg.joiner_dim, dd, 1.2)
    put("decoder.decoder_proj.weight", w); put("decoder.decoder_proj.bias", b)

    # ---- joiner: peaky logits, dominant blank
    w, b = _linear(rng, cfg.vocab_size, cfg.joiner_dim, 7.0, bias_std=0.5)
    sigma = 7.0 * 0.62
    b[cfg.blank_id] = sigma * (math.sqrt(2.0 * math.log(cfg.vocab_size)) + 1.35)
    put("joiner.output_linear.weight", w); put("joiner.output_linear.bias", b)
    return W


# --------------------------------------------------------------------------- container IO
def config_items(cfg: ZipformerConfig):
    tup = lambda t: ",".join(str(int(v)) for v in t)
    return [("name", cfg.name), ("num_encoder_layers", tup(cfg.num_encoder_layers)),
            ("downsampling_factor", tup(cfg.downsampling_factor)), ("encoder_dim", tup(cfg.encoder_dim)),
            ("feedforward_dim", tup(cfg.feedforward_dim)), ("num_heads", tup(cfg.num_heads)),
            ("cnn_module_kernel", tup(cfg.cnn_module_kernel)), ("query_head_dim", cfg.query_head_dim),
            ("pos_head_dim", cfg.pos_head_dim), ("value_head_dim", cfg.value_head_dim),
            ("pos_dim", cfg.pos_dim), ("feature_dim", cfg.feature_dim), ("decoder_dim", cfg.decoder_dim),
            ("joiner_dim", cfg.joiner_dim), ("context_size", cfg.context_size),
            ("vocab_size", cfg.vocab_size), ("blank_id", cfg.blank_id), ("unk_id", cfg.unk_id)]


def save_container(path: str, cfg: ZipformerConfig, tensors: dict, prefix: str = "") -> None:
    names = [n for n in tensors if n.startswith(prefix)]
    offs, off = {}, 0
    for n in names:
        off = (off + 255) // 256 * 256
        offs[n] = off
        off += tensors[n].nbytes
    body = "".join(f"config {k} {v}\n" for k, v in (config_items(cfg) if cfg is not None else [("name", "vad")]))
    for n in names:
        a = tensors[n]
        dims = " ".join(str(d) for d in a.shape)
        body += f"tensor {n} f32 {a.ndim} {dims} {offs[n]} {a.nbytes}\n"
    body += "end\n"
    hb = 4096
    while True:
        first = f"{MAGIC} 1 {hb}\n"
        if len(first.encode()) + len(body.encode()) <= hb:
            break
        hb += 4096
    header = (first + body).encode()
    header += b"\0" * (hb - len(header))
    with open(path, "wb") as f:
        f.write(header)
        pos = 0
        for n in names:
            f.write(b"\0" * (offs[n] - pos))
            f.write(tensors[n].tobytes())
            pos = offs[n] + tensors[n].nbytes


def stft_basis() -> np.ndarray:
    """[258, 256] Hann-windowed Fourier basis of the VAD front end: rows 0..128 real parts, 129..257 imaginary parts."""
    n = np.arange(256, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / 256.0)
    ang = 2.0 * np.pi * np.arange(129, dtype=np.float64)[:, None] * n[None, :] / 256.0
    return np.concatenate([np.cos(ang) * hann, -np.sin(ang) * hann], axis=0).astype(np.float32)


def init_vad_weights(seed: int = 5) -> dict:
    """Seeded random tensors of the voice-activity network (Silero VAD v5 topology, 16 kHz; the reference loads
    `silero_vad.onnx`, core/vad_utils.py:31-52, which is not available offline): STFT basis, four Conv1d, LSTMCell(128, 128),
    Conv1d(128 -> 1). Scales keep activations O(1) and the recurrence contractive."""
    rng = np.random.default_rng(seed)
    W = {"vad.stft.basis": stft_basis()}

    def conv(name, co, ci, k, gain):
        W[name + ".weight"] = (rng.standard_normal((co, ci, k)) * (gain / math.sqrt(ci * k))).astype(np.float32)
        W[name + ".bias"] = (rng.standard_normal(co) * 0.05).astype(np.float32)

    conv("vad.enc0", 128, 129, 3, 0.6)
    conv("vad.enc1", 64, 128, 3, 1.4)
    conv("vad.enc2", 64, 64, 3, 1.4)
    conv("vad.enc3", 128, 64, 3, 1.4)
    W["vad.lstm.weight_ih"] = (rng.standard_normal((512, 128)) * (1.0 / math.sqrt(128))).astype(np.float32)
    W["vad.lstm.weight_hh"] = (rng.standard_normal((512, 128)) * (0.6 / math.sqrt(128))).astype(np.float32)
    W["vad.lstm.bias_ih"] = (rng.standard_normal(512) * 0.05).astype(np.float32)
    W["vad.lstm.bias_hh"] = (rng.standard_normal(512) * 0.05).astype(np.float32)
    W["vad.out.weight"] = (rng.standard_normal((1, 128, 1)) * 0.8).astype(np.float32)
    W["vad.out.bias"] = np.array([-0.3], dtype=np.float32)
    return W


def save_vad(path: str, tensors: dict) -> str:
    """Writes the vad.* tensors as a `.b200w` container (no model config lines)."""
    save_container(path, None, tensors, prefix="vad.")
    return path


def load_container(path: str):
    """Returns (config dict of str->str, tensors dict)."""
    with open(path, "rb") as f:
        first = f.readline().decode()
        magic, ver, hb = first.split()
        if magic != MAGIC:
            raise ValueError(f"{path}: not a {MAGIC} container")
        hb = int(hb)
        f.seek(0)
        header = f.read(hb).split(b"\0", 1)[0].decode()
        blob = np.frombuffer(f.read(), dtype=np.uint8)
    cfg, tensors = {}, {}
    for line in header.splitlines()[1:]:
        parts = line.split()
        if not parts or parts[0] == "end":
            break
        if parts[0] == "config":
            cfg[parts[1]] = parts[2] if len(parts) > 2 else ""
        elif parts[0] == "tensor":
            nd = int(parts[3])
            dims = tuple(int(x) for x in parts[4:4 + nd])
            o, nb = int(parts[4 + nd]), int(parts[5 + nd])
            tensors[parts[1]] = blob[o:o + nb].view(np.float32).reshape(dims)
    return cfg, tensors


def config_from_dict(d: dict) -> ZipformerConfig:
    t = lambda s: tuple(int(x) for x in s.split(","))
    return ZipformerConfig(
        name=d.get("name", "zipformer"), num_encoder_layers=t(d["num_encoder_layers"]),
        downsampling_factor=t(d["downsampling_factor"]), encoder_dim=t(d["encoder_dim"]),
        feedforward_dim=t(d["feedforward_dim"]), num_heads=t(d["num_heads"]),
        cnn_module_kernel=t(d["cnn_module_kernel"]), query_head_dim=int(d["query_head_dim"]),
        pos_head_dim=int(d["pos_head_dim"]), value_head_dim=int(d["value_head_dim"]),
        pos_dim=int(d["pos_dim"]), feature_dim=int(d["feature_dim"]), decoder_dim=int(d["decoder_dim"]),
        joiner_dim=int(d["joiner_dim"]), context_size=int(d["context_size"]),
        vocab_size=int(d["vocab_size"]), blank_id=int(d["blank_id"]), unk_id=int(d["unk_id"]))


def make_tokens(cfg: ZipformerConfig, seed: int = 7):
    """Synthetic upper-case BPE-like vocabulary in sherpa's `tokens.txt` layout (`<sym> <id>`).

    ids 0/1/2 = <blk>/<sos/eos>/<unk> (reference core/asr_engine.py:1034-1035); ~45 % of the
    pieces start a word with U+2581, as in the reference's BPE-2000 vocabulary.
    """
    rng = np.random.default_rng(seed)
    syll = ["A", "E", "I", "O", "U", "Y", "NG", "NH", "TH", "TR", "CH", "KH", "PH", "B", "C", "D",
            "G", "H", "K", "L", "M", "N", "P", "Q", "R", "S", "T", "V", "X", "Ô", "Ơ", "Ư", "Â", "Ê", "Đ"]
    toks, seen = ["<blk>", "<sos/eos>", "<unk>"], set()
    while len(toks) < cfg.vocab_size:
        n = int(rng.integers(1, 4))
        s = "".join(syll[int(j)] for j in rng.integers(0, len(syll), n))
        if rng.random() < 0.45:
            s = "▁" + s
        if s in seen:
            continue
        seen.add(s)
        toks.append(s)
    return toks


def write_model_dir(model_dir: str, cfg: ZipformerConfig, seed: int) -> dict:
    """Writes encoder-/decoder-/joiner- containers + tokens.txt the way a sherpa model dir is laid
    out (reference core/asr_engine.py:912-927 looks for `encoder-*`, `decoder-*`, `joiner-*`, `tokens.txt`)."""
    os.makedirs(model_dir, exist_ok=True)
    W = init_weights(cfg, seed)
    paths = {}
    for part in ("encoder", "decoder", "joiner"):
        paths[part] = os.path.join(model_dir, f"{part}-{cfg.name}.b200w")
        save_container(paths[part], cfg, W, prefix=part + ".")
    paths["tokens"] = os.path.join(model_dir, "tokens.txt")
    with open(paths["tokens"], "w", encoding="utf-8") as f:
        for i, t in enumerate(make_tokens(cfg)):
            f.write(f"{t} {i}\n")
    return paths
