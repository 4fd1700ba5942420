"""ORACLE (test infrastructure, never shipped): Zipformer2 encoder / stateless decoder / joiner on the CPU.

The reference executes three ONNX graphs exported from k2-fsa/icefall `egs/librispeech/ASR/zipformer`
(`zipformer.py`, `subsampling.py`, `scaling.py`, `decoder.py`, `joiner.py`, `export-onnx.py`; not vendored,
model files absent) through onnxruntime (/root/reference core/asr_engine.py:1045-1056,1084-1093).
This file restates the published inference arithmetic of that architecture in plain
PyTorch, batch 1 as the reference always runs it (core/asr_engine.py:1045-1047), in fp32 or fp64.

Parity status: "parity unpinned" against the real ONNX graphs (no checkpoint, no onnxruntime offline).
It is pinned structurally by the parameter-count cross-check against the manifest byte sizes
(tests/test_oracle_model.py) and numerically by its own fp64 run.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def swoosh_l(x):
    return torch.logaddexp(torch.zeros((), dtype=x.dtype), x - 4.0) - 0.08 * x - 0.035


def swoosh_r(x):
    return torch.logaddexp(torch.zeros((), dtype=x.dtype), x - 1.0) - 0.08 * x - 0.313261687


def bias_norm(x, bias, log_scale):
    """icefall scaling.py BiasNorm: x * mean((x-bias)^2)^-0.5 * exp(log_scale)."""
    scales = (torch.mean((x - bias) ** 2, dim=-1, keepdim=True) ** -0.5) * log_scale.exp()
    return x * scales


class Weights:
    """Name -> torch tensor view over the container's numpy tensors."""

    def __init__(self, tensors: dict, dtype=torch.float32):
        self.t = {k: torch.from_numpy(np.array(v, copy=True)).to(dtype) for k, v in tensors.items()}
        self.dtype = dtype

    def __getitem__(self, k):
        return self.t[k]

    def lin(self, x, name):
        return F.linear(x, self.t[name + ".weight"], self.t.get(name + ".bias"))


# ----------------------------------------------------------------------------- encoder_embed
def encoder_embed(W: Weights, x):
    """icefall subsampling.py Conv2dSubsampling. x[T,80] -> [T1,D0]."""
    e = "encoder.embed."
    x = x[None, None]                                                  # [1,1,T,80]
    x = swoosh_r(F.conv2d(x, W[e + "conv0.weight"], W[e + "conv0.bias"], padding=(0, 1)))
    x = swoosh_r(F.conv2d(x, W[e + "conv1.weight"], W[e + "conv1.bias"], stride=2))
    x = swoosh_r(F.conv2d(x, W[e + "conv2.weight"], W[e + "conv2.bias"], stride=(1, 2)))
    # ConvNeXt: x + pw2(SwooshL(pw1(dw7x7(x))))
    y = F.conv2d(x, W[e + "convnext.dw.weight"], W[e + "convnext.dw.bias"], padding=3, groups=x.shape[1])
    y = F.conv2d(y, W[e + "convnext.pw1.weight"][:, :, None, None], W[e + "convnext.pw1.bias"])
    y = swoosh_l(y)
    y = F.conv2d(y, W[e + "convnext.pw2.weight"][:, :, None, None], W[e + "convnext.pw2.bias"])
    x = x + y                                                          # [1,128,T1,19]
    b, c, t, f = x.shape
    x = x.transpose(1, 2).reshape(t, c * f)                            # channel-major
    x = W.lin(x, e + "out")
    return bias_norm(x, W[e + "out_norm.bias"], W[e + "out_norm.log_scale"])


# ----------------------------------------------------------------------------- stack pieces
def compact_rel_pos(Tk: int, pos_dim: int, dtype):
    """icefall zipformer.py CompactRelPositionalEncoding -> [2Tk-1, pos_dim]."""
    x = torch.arange(-(Tk - 1), Tk, dtype=torch.float32).unsqueeze(1)
    freqs = 1 + torch.arange(pos_dim // 2, dtype=torch.float32)
    cl = pos_dim ** 0.5
    xc = cl * x.sign() * ((x.abs() + cl).log() - math.log(cl))
    length_scale = pos_dim / (2.0 * math.pi)
    x_atan = (xc / length_scale).atan()
    pe = torch.zeros(x.shape[0], pos_dim, dtype=torch.float32)
    pe[:, 0::2] = (x_atan * freqs).cos()
    pe[:, 1::2] = (x_atan * freqs).sin()
    pe[:, -1] = 1.0
    return pe.to(dtype)


def simple_downsample(x, bias, ds: int):
    """icefall SimpleDownsample: right-pad by repeating the last frame, softmax(bias)-weighted sum."""
    T, C = x.shape
    d = (T + ds - 1) // ds
    pad = d * ds - T
    if pad:
        x = torch.cat([x, x[-1:].expand(pad, C)], dim=0)
    w = bias.softmax(dim=0)
    return (x.reshape(d, ds, C) * w[None, :, None]).sum(dim=1)


def attn_weights(W: Weights, p: str, x, pos_emb, H: int, qd: int, pd: int):
    """RelPositionMultiheadAttentionWeights (no 1/sqrt(d): baked into in_proj). -> [H,Tk,Tk]."""
    Tk = x.shape[0]
    proj = W.lin(x, p + "attn_w.in_proj")
    q = proj[:, : H * qd].reshape(Tk, H, qd).permute(1, 0, 2)
    k = proj[:, H * qd: 2 * H * qd].reshape(Tk, H, qd).permute(1, 2, 0)
    pp = proj[:, 2 * H * qd:].reshape(Tk, H, pd).permute(1, 0, 2)
    scores = torch.matmul(q, k)                                         # [H,Tk,Tk]
    pos = F.linear(pos_emb, W[p + "attn_w.linear_pos.weight"])          # [2Tk-1, H*pd]
    pos = pos.reshape(2 * Tk - 1, H, pd).permute(1, 2, 0)               # [H,pd,2Tk-1]
    ps = torch.matmul(pp, pos)                                          # [H,Tk,2Tk-1]
    i = torch.arange(Tk)[:, None]
    j = torch.arange(Tk)[None, :]
    idx = (Tk - 1) - i + j
    ps = torch.gather(ps, 2, idx[None].expand(H, Tk, Tk))
    return (scores + ps).softmax(dim=-1)


def self_attn(W: Weights, p: str, x, A, H: int, vd: int):
    Tk = x.shape[0]
    v = W.lin(x, p + ".in").reshape(Tk, H, vd).permute(1, 0, 2)
    o = torch.matmul(A, v).permute(1, 0, 2).reshape(Tk, H * vd)
    return W.lin(o, p + ".out")


def nonlin_attn(W: Weights, p: str, x, A0):
    h3 = W.lin(x, p + "nonlin.in")
    s, xx, y = h3.chunk(3, dim=-1)
    xx = xx * torch.tanh(s)
    xx = torch.matmul(A0, xx)
    xx = xx * y
    return W.lin(xx, p + "nonlin.out")


def conv_module(W: Weights, p: str, x, k: int):
    D = x.shape[1]
    h = W.lin(x, p + ".in")
    xx, s = h.chunk(2, dim=-1)
    xx = xx * torch.sigmoid(s)
    xx = xx.t()[None]                                                   # [1,D,T]
    xx = F.conv1d(xx, W[p + ".dw.weight"], W[p + ".dw.bias"], padding=k // 2, groups=D)
    xx = xx[0].t()
    return W.lin(swoosh_r(xx), p + ".out")


def feed_forward(W: Weights, p: str, x):
    return W.lin(swoosh_l(W.lin(x, p + ".in")), p + ".out")


def encoder_layer(W: Weights, p: str, x, pos_emb, H, k, cfg):
    orig = x
    A = attn_weights(W, p, x, pos_emb, H, cfg.query_head_dim, cfg.pos_head_dim)
    x = x + feed_forward(W, p + "ff1", x)
    x = x + nonlin_attn(W, p, x, A[0])
    x = x + self_attn(W, p + "attn1", x, A, H, cfg.value_head_dim)
    x = x + conv_module(W, p + "conv1", x, k)
    x = x + feed_forward(W, p + "ff2", x)
    x = orig + (x - orig) * W[p + "bypass_mid.scale"]
    x = x + self_attn(W, p + "attn2", x, A, H, cfg.value_head_dim)
    x = x + conv_module(W, p + "conv2", x, k)
    x = x + feed_forward(W, p + "ff3", x)
    x = bias_norm(x, W[p + "norm.bias"], W[p + "norm.log_scale"])
    return orig + (x - orig) * W[p + "bypass.scale"]


def convert_channels(x, D: int):
    c = x.shape[1]
    if D <= c:
        return x[:, :D]
    return torch.cat([x, torch.zeros(x.shape[0], D - c, dtype=x.dtype)], dim=1)


def encoder(W: Weights, cfg, feats, return_intermediate=False):
    """feats[T,80] -> encoder_out[T',joiner_dim] (encoder_proj included, as in the ONNX export).
    T1 = (T-7)//2, T' = (T1+1)//2."""
    x = torch.as_tensor(feats).to(W.dtype)
    inter = {}
    x = encoder_embed(W, x)
    inter["embed"] = x
    outs = []
    for i, (L, ds, D, H, k) in enumerate(zip(cfg.num_encoder_layers, cfg.downsampling_factor,
                                             cfg.encoder_dim, cfg.num_heads, cfg.cnn_module_kernel)):
        s = f"encoder.stack{i}."
        x = convert_channels(x, D)
        src_orig = x
        if ds > 1:
            x = simple_downsample(x, W[s + "downsample.bias"], ds)
        pos_emb = compact_rel_pos(x.shape[0], cfg.pos_dim, W.dtype)
        for l in range(L):
            x = encoder_layer(W, s + f"layer{l}.", x, pos_emb, H, k, cfg)
        if ds > 1:
            x = x.repeat_interleave(ds, dim=0)[: src_orig.shape[0]]
            x = src_orig + (x - src_orig) * W[s + "out_combiner.scale"]
        outs.append(x)
        inter[f"stack{i}"] = x
    pieces, cur = [outs[-1]], cfg.encoder_dim[-1]
    for i in range(len(outs) - 2, -1, -1):
        d = cfg.encoder_dim[i]
        if d > cur:
            pieces.append(outs[i][:, cur:d])
            cur = d
    x = torch.cat(pieces, dim=1)
    x = simple_downsample(x, W["encoder.downsample_output.bias"], 2)
    x = W.lin(x, "encoder.encoder_proj")
    if return_intermediate:
        return x, inter
    return x


# ----------------------------------------------------------------------------- decoder / joiner
def decoder(W: Weights, cfg, y):
    """y[M,2] int64 (already max(0,.) per core/asr_engine.py:1052,1075) -> [M,joiner_dim]."""
    y = torch.as_tensor(y, dtype=torch.int64)
    emb = W["decoder.embedding.weight"][y.clamp(min=0)] * (y >= 0).unsqueeze(-1).to(W.dtype)
    e = F.conv1d(emb.permute(0, 2, 1), W["decoder.conv.weight"], None, groups=cfg.decoder_dim // 4)
    e = F.relu(e.permute(0, 2, 1)).squeeze(1)
    return W.lin(e, "decoder.decoder_proj")


def joiner(W: Weights, enc, dec):
    return W.lin(torch.tanh(enc + dec), "joiner.output_linear")


# ----------------------------------------------------------------------------- ORT-like sessions
class _Out:
    def __init__(self, shape):
        self.shape = shape


class EncoderSession:
    """Duck-types `InferenceSession.run(None, {"x","x_lens"})` (core/asr_engine.py:1047)."""

    def __init__(self, W, cfg):
        self.W, self.cfg = W, cfg

    def run(self, _, feeds):
        x, x_lens = feeds["x"], feeds["x_lens"]
        outs, lens = [], []
        with torch.no_grad():
            for n in range(x.shape[0]):
                o = encoder(self.W, self.cfg, x[n, : int(x_lens[n])])
                outs.append(o.to(torch.float32).numpy())
                lens.append(o.shape[0])
        Tm = max(lens)
        out = np.zeros((len(outs), Tm, outs[0].shape[1]), dtype=np.float32)
        for n, o in enumerate(outs):
            out[n, : o.shape[0]] = o
        return [out, np.array(lens, dtype=np.int64)]


class DecoderSession:
    def __init__(self, W, cfg, batched=False):
        self.W, self.cfg, self.batched = W, cfg, batched

    def run(self, _, feeds):
        if self.batched:      # one call for all rows, as onnxruntime runs it (timing runs; results may differ in the last bits)
            with torch.no_grad():
                return [decoder(self.W, self.cfg, np.asarray(feeds["y"])).to(torch.float32).numpy()]
        # row by row: BLAS results depend on the batch shape in the last bits, and the reference's memo
        # (core/asr_engine.py:1072-1088) batches whatever contexts happen to miss; per-row evaluation makes
        # a context's decoder output a pure function of the context.
        y = np.asarray(feeds["y"])
        with torch.no_grad():
            rows = [decoder(self.W, self.cfg, y[i:i + 1]).to(torch.float32).numpy() for i in range(y.shape[0])]
        return [np.concatenate(rows, axis=0)]


class JoinerSession:
    def __init__(self, W, cfg, batched=False):
        self.W, self.cfg, self.batched = W, cfg, batched

    def get_outputs(self):
        return [_Out(["N", self.cfg.vocab_size])]

    def run(self, _, feeds):
        with torch.no_grad():
            e = torch.from_numpy(np.ascontiguousarray(feeds["encoder_out"])).to(self.W.dtype)
            d = torch.from_numpy(np.ascontiguousarray(feeds["decoder_out"])).to(self.W.dtype)
            if self.batched:
                return [joiner(self.W, e, d).to(torch.float32).numpy()]
            rows = [joiner(self.W, e[i:i + 1], d[i:i + 1]).to(torch.float32).numpy() for i in range(e.shape[0])]
            return [np.concatenate(rows, axis=0)]


def make_recognizer(tensors: dict, cfg, id2token=None, max_active_paths=4, context_graph=None,
                    dtype=torch.float32, batched_rows=False):
    """Builds the dict `create_recognizer` returns (core/asr_engine.py:1005-1012) over oracle sessions. batched_rows: the
    decoder / joiner sessions evaluate all rows of a call at once like onnxruntime does (the CPU timing arm); the parity
    tests keep the row-by-row default, which makes every row a pure function of its inputs."""
    W = Weights(tensors, dtype)
    return {
        "enc_sess": EncoderSession(W, cfg), "dec_sess": DecoderSession(W, cfg, batched_rows),
        "joi_sess": JoinerSession(W, cfg, batched_rows), "id2token": id2token or {}, "vocab_size": cfg.vocab_size,
        "max_active_paths": max_active_paths, "model_path": "", "dec_cache": {},
        "context_graph": context_graph, "provider_info": {},
    }


def count_params(tensors: dict, prefix: str) -> int:
    return int(sum(v.size for k, v in tensors.items() if k.startswith(prefix)))
