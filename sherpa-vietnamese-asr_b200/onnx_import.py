"""ONNX initialisers -> `.b200w` containers (the reference loads `encoder-*.onnx`, `decoder-*.onnx`, `joiner-*.onnx` per model
directory, /root/reference core/asr_engine.py:912-927; the engine loads its own containers, weights.py).

STATUS - read before relying on it: neither the `onnx` package nor a single real checkpoint is available offline, so this
converter has been exercised ONLY on synthetic files written by `write_model` below in the naming icefall's `export-onnx.py`
is known to produce (module-path initialiser names for convolutions, biases and norm parameters; Linear weights folded to
anonymous transposed `onnx::MatMul_<n>` initialisers that are recognised through the MatMul node's name or the bias of the Add
that consumes it). It has never seen a real export. It therefore verifies what it can and refuses the rest: every tensor the
engine needs must be found with the expected shape, the architecture is inferred from the shapes and checked for consistency,
and anything unmapped or missing is reported by name instead of being guessed. Export-time folds that leave no named trace
(e.g. `exp(log_scale)` of a BiasNorm merged into a constant) are NOT reconstructed - such a model is rejected with the names of
the missing tensors.

No third-party dependency: a minimal protobuf wire-format reader for the four message types involved (ModelProto.graph = 7,
GraphProto.node = 1 / initializer = 5, NodeProto input/output/name/op_type = 1/2/3/4, TensorProto dims/data_type/float_data/
int64_data/name/raw_data = 1/2/4/7/8/9).
"""
from __future__ import annotations

import glob
import os
import re
import shutil
import struct
from typing import Dict, List, Tuple

import numpy as np

from . import weights


# --------------------------------------------------------------------------- protobuf wire format (subset)
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yields (field number, wire type, value) of one message; length-delimited values as memoryview slices."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


_DTYPES = {1: np.float32, 7: np.int64, 6: np.int32, 11: np.float64, 10: np.float16}


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[float] = []
    ints: List[int] = []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            if wt == 0:
                dims.append(v)
            else:                                   # packed
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    dims.append(d)
        elif fn == 2:
            dtype = v
        elif fn == 4:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", v) if wt == 2 else struct.unpack("<f", v))
        elif fn == 7:
            if wt == 0:
                ints.append(v)
            else:
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    ints.append(d)
        elif fn == 8:
            name = bytes(v).decode("utf-8")
        elif fn == 9:
            raw = bytes(v)
    if dtype not in _DTYPES:
        return name, None
    if raw is not None:
        arr = np.frombuffer(raw, dtype=_DTYPES[dtype]).copy()
    elif floats:
        arr = np.asarray(floats, dtype=np.float32)
    else:
        arr = np.asarray(ints, dtype=np.int64)
    return name, arr.reshape(dims) if dims else arr.reshape(())


def read_model(path: str) -> dict:
    """-> {"initializers": {name: ndarray}, "nodes": [{"name", "op_type", "inputs", "outputs"}]}"""
    with open(path, "rb") as f:
        buf = f.read()
    graph = None
    for fn, wt, v in _fields(buf):
        if fn == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError(f"{path}: no GraphProto (not an ONNX ModelProto?)")
    inits: Dict[str, np.ndarray] = {}
    nodes = []
    for fn, wt, v in _fields(graph):
        if fn == 5 and wt == 2:
            name, arr = _tensor(v)
            if arr is not None:
                inits[name] = arr
        elif fn == 1 and wt == 2:
            node = {"name": "", "op_type": "", "inputs": [], "outputs": []}
            for f2, w2, v2 in _fields(v):
                if w2 != 2:
                    continue
                if f2 == 1:
                    node["inputs"].append(bytes(v2).decode("utf-8"))
                elif f2 == 2:
                    node["outputs"].append(bytes(v2).decode("utf-8"))
                elif f2 == 3:
                    node["name"] = bytes(v2).decode("utf-8")
                elif f2 == 4:
                    node["op_type"] = bytes(v2).decode("utf-8")
            nodes.append(node)
    return {"initializers": inits, "nodes": nodes}


# --------------------------------------------------------------------------- writer (tests; documents the expected shape of a file)
def _enc_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(fn: int, payload: bytes) -> bytes:
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def write_model(path: str, initializers: Dict[str, np.ndarray], nodes: List[dict]) -> None:
    g = bytearray()
    for nd in nodes:
        m = b"".join(_ld(1, s.encode()) for s in nd.get("inputs", [])) + b"".join(_ld(2, s.encode()) for s in nd.get("outputs", []))
        m += _ld(3, nd.get("name", "").encode()) + _ld(4, nd["op_type"].encode())
        g += _ld(1, m)
    for name, arr in initializers.items():
        a = np.ascontiguousarray(arr)
        dt = {np.dtype(np.float32): 1, np.dtype(np.int64): 7}[a.dtype]
        t = b"".join(_enc_varint((1 << 3) | 0) + _enc_varint(int(d)) for d in a.shape)
        t += _enc_varint((2 << 3) | 0) + _enc_varint(dt) + _ld(8, name.encode()) + _ld(9, a.tobytes())
        g += _ld(5, t)
    with open(path, "wb") as f:
        f.write(_enc_varint((1 << 3) | 0) + _enc_varint(8) + _ld(7, bytes(g)))          # ir_version, graph


# --------------------------------------------------------------------------- icefall module paths -> container names
_LAYER = [
    (r"self_attn_weights\.in_proj", "attn_w.in_proj"), (r"self_attn_weights\.linear_pos", "attn_w.linear_pos"),
    (r"feed_forward([123])\.in_proj", r"ff\1.in"), (r"feed_forward([123])\.out_proj", r"ff\1.out"),
    (r"nonlin_attention\.in_proj", "nonlin.in"), (r"nonlin_attention\.out_proj", "nonlin.out"),
    (r"self_attn([12])\.in_proj", r"attn\1.in"), (r"self_attn([12])\.out_proj", r"attn\1.out"),
    (r"conv_module([12])\.in_proj", r"conv\1.in"), (r"conv_module([12])\.depthwise_conv", r"conv\1.dw"),
    (r"conv_module([12])\.out_proj", r"conv\1.out"), (r"norm", "norm"),
]
_EMBED = {"conv.0": "conv0", "conv.4": "conv1", "conv.7": "conv2", "convnext.depthwise_conv": "convnext.dw",
          "convnext.pointwise_conv1": "convnext.pw1", "convnext.pointwise_conv2": "convnext.pw2", "out": "out", "out_norm": "out_norm"}


def container_name(path: str):
    """icefall parameter path (`encoder.encoders.2.encoder.layers.1.feed_forward1.in_proj.weight`, `encoder_embed.conv.4.bias`,
    `decoder_proj.weight` ...) -> the engine's tensor name, or None when the parameter is not one the engine reads."""
    m = re.fullmatch(r"encoder_embed\.(.+)\.(weight|bias|log_scale)", path)
    if m and m.group(1) in _EMBED:
        return f"encoder.embed.{_EMBED[m.group(1)]}.{m.group(2)}"
    m = re.fullmatch(r"encoder\.encoders\.(\d+)\.(?:encoder\.)?layers\.(\d+)\.(.+)", path)
    if m:
        i, l, rest = m.group(1), m.group(2), m.group(3)
        base = f"encoder.stack{i}.layer{l}."
        mm = re.fullmatch(r"(bypass|bypass_mid)\.bypass_scale", rest)
        if mm:
            return base + mm.group(1) + ".scale"
        if "." not in rest:
            return None
        module, param = rest.rsplit(".", 1)
        if param not in ("weight", "bias", "log_scale"):
            return None
        for pat, rep in _LAYER:
            mm = re.fullmatch(pat, module)
            if mm:
                return base + mm.expand(rep) + "." + param
        return None
    m = re.fullmatch(r"encoder\.encoders\.(\d+)\.downsample\.bias", path)
    if m:
        return f"encoder.stack{m.group(1)}.downsample.bias"
    m = re.fullmatch(r"encoder\.encoders\.(\d+)\.out_combiner\.bypass_scale", path)
    if m:
        return f"encoder.stack{m.group(1)}.out_combiner.scale"
    fixed = {"encoder.downsample_output.bias": "encoder.downsample_output.bias",
             "encoder_proj.weight": "encoder.encoder_proj.weight", "encoder_proj.bias": "encoder.encoder_proj.bias",
             "decoder.embedding.weight": "decoder.embedding.weight", "decoder.conv.weight": "decoder.conv.weight",
             "decoder_proj.weight": "decoder.decoder_proj.weight", "decoder_proj.bias": "decoder.decoder_proj.bias",
             "output_linear.weight": "joiner.output_linear.weight", "output_linear.bias": "joiner.output_linear.bias"}
    return fixed.get(path)


def _module_of_node(node_name: str) -> str:
    """`/encoder/encoders.0/layers.1/feed_forward1/in_proj/MatMul` -> `encoder.encoders.0.layers.1.feed_forward1.in_proj`"""
    parts = [p for p in node_name.split("/") if p]
    if parts and re.fullmatch(r"(MatMul|Gemm|Add|Conv)(_\d+)?", parts[-1]):
        parts = parts[:-1]
    return ".".join(parts)


def named_parameters(model: dict) -> Tuple[Dict[str, np.ndarray], List[str]]:
    """Initialisers under their icefall parameter paths: named ones as they are; anonymous Linear weights (`onnx::MatMul_*`,
    stored transposed [in, out]) through the MatMul node that reads them - its name, or the bias of the Add that consumes its
    output. Returns (parameters, anonymous initialisers that could not be placed)."""
    inits, nodes = model["initializers"], model["nodes"]
    params = {n: a for n, a in inits.items() if not n.startswith("onnx::") and not re.fullmatch(r"\d+", n)}
    consumer_bias = {}
    for nd in nodes:
        if nd["op_type"] == "Add":
            named = [i for i in nd["inputs"] if i in params and i.endswith(".bias")]
            other = [i for i in nd["inputs"] if i not in inits]
            if len(named) == 1 and len(other) == 1:
                consumer_bias[other[0]] = named[0]
    placed = set()
    for nd in nodes:
        if nd["op_type"] != "MatMul":
            continue
        w = [i for i in nd["inputs"] if i in inits and i not in params]
        if len(w) != 1 or inits[w[0]].ndim != 2:
            continue
        module = None
        out = nd["outputs"][0] if nd["outputs"] else None
        if out in consumer_bias:
            module = consumer_bias[out][: -len(".bias")]
        elif nd["name"]:
            module = _module_of_node(nd["name"])
        if module:
            params[module + ".weight"] = np.ascontiguousarray(inits[w[0]].T)      # [in, out] -> [out, in]
            placed.add(w[0])
    # Linear layers on 2-D inputs are exported as Gemm(A, B, C): B anonymous -> the module comes from the named bias C (or the
    # node name); whether B is [out, in] (transB = 1, the usual export) or [in, out] is read off the bias length
    for nd in nodes:
        if nd["op_type"] != "Gemm" or len(nd["inputs"]) < 2:
            continue
        b = nd["inputs"][1]
        if b not in inits or b in params or inits[b].ndim != 2:
            continue
        c = nd["inputs"][2] if len(nd["inputs"]) > 2 else None
        module = c[: -len(".bias")] if c in params and c.endswith(".bias") else (_module_of_node(nd["name"]) if nd["name"] else None)
        if not module:
            continue
        w = inits[b]
        if c in params and w.shape[0] != params[c].shape[0] and w.shape[1] == params[c].shape[0]:
            w = w.T
        params[module + ".weight"] = np.ascontiguousarray(w)
        placed.add(b)
    unplaced = [n for n, a in inits.items() if n not in params and n not in placed and getattr(a, "ndim", 0) >= 2]
    return params, unplaced


# --------------------------------------------------------------------------- conversion
def _shape_fix(name: str, a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.float32)
    if name.endswith(("convnext.pw1.weight", "convnext.pw2.weight")) and a.ndim == 4:       # Conv2d 1x1 -> Linear
        a = a.reshape(a.shape[0], a.shape[1])
    if name.endswith("log_scale"):
        a = a.reshape(1)
    if name.endswith("downsample.bias") or name.endswith("downsample_output.bias"):
        a = a.reshape(-1)
    return np.ascontiguousarray(a)


def infer_config(t: Dict[str, np.ndarray], name: str) -> weights.ZipformerConfig:
    """Architecture from the tensor shapes; raises if a stack is inconsistent."""
    n_stacks = 1 + max(int(m.group(1)) for m in (re.match(r"encoder\.stack(\d+)\.", k) for k in t) if m)
    L, ds, D, F, H, K = [], [], [], [], [], []
    qd = pd = vd = None
    for i in range(n_stacks):
        p = f"encoder.stack{i}."
        L.append(1 + max(int(m.group(1)) for m in (re.match(re.escape(p) + r"layer(\d+)\.", k) for k in t) if m))
        D.append(int(t[p + "layer0.norm.bias"].shape[0]))
        F.append(int(t[p + "layer0.ff2.in.weight"].shape[0]))
        K.append(int(t[p + "layer0.conv1.dw.weight"].shape[-1]))
        ds.append(int(t[p + "downsample.bias"].shape[0]) if p + "downsample.bias" in t else 1)
        hp = int(t[p + "layer0.attn_w.linear_pos.weight"].shape[0])          # H * pos_head_dim
        hv = int(t[p + "layer0.attn1.in.weight"].shape[0])                   # H * value_head_dim
        hq = int(t[p + "layer0.attn_w.in_proj.weight"].shape[0]) - hp        # H * 2 * query_head_dim
        if pd is None:
            # heads are not stored: the published configs use pos_head_dim 4, which fixes H and with it the other two
            pd = 4
        h = hp // pd
        if h <= 0 or hp % pd or hv % h or hq % (2 * h):
            raise ValueError(f"stack {i}: attention projections do not factor into heads (H*pd={hp}, H*vd={hv}, 2*H*qd={hq})")
        if qd is None:
            qd, vd = hq // (2 * h), hv // h
        elif (qd, vd) != (hq // (2 * h), hv // h):
            raise ValueError(f"stack {i}: head dims differ from stack 0")
        H.append(h)
    return weights.ZipformerConfig(
        name=name, num_encoder_layers=tuple(L), downsampling_factor=tuple(ds), encoder_dim=tuple(D), feedforward_dim=tuple(F),
        num_heads=tuple(H), cnn_module_kernel=tuple(K), query_head_dim=qd, pos_head_dim=pd, value_head_dim=vd,
        pos_dim=int(t["encoder.stack0.layer0.attn_w.linear_pos.weight"].shape[1]), feature_dim=80,
        decoder_dim=int(t["decoder.embedding.weight"].shape[1]), joiner_dim=int(t["joiner.output_linear.weight"].shape[1]),
        context_size=int(t["decoder.conv.weight"].shape[2]), vocab_size=int(t["decoder.embedding.weight"].shape[0]))


def convert_model_dir(src_dir: str, dst_dir: str, name: str = "zipformer-onnx") -> dict:
    """`encoder-*.onnx`, `decoder-*.onnx`, `joiner-*.onnx` (+ tokens.txt) in src_dir -> the three `.b200w` containers in dst_dir.
    Returns {"config", "paths", "unmapped", "unplaced"}; raises ValueError listing every tensor the engine needs that was not
    found or has the wrong shape."""
    tensors: Dict[str, np.ndarray] = {}
    unmapped, unplaced = [], []
    for part in ("encoder", "decoder", "joiner"):
        files = sorted(f for f in glob.glob(os.path.join(src_dir, part + "-*.onnx")) if ".int8." not in f) or \
            sorted(glob.glob(os.path.join(src_dir, part + "-*.onnx")))
        if not files:
            raise FileNotFoundError(f"no {part}-*.onnx in {src_dir}")
        params, up = named_parameters(read_model(files[0]))
        unplaced += [f"{part}: {n}" for n in up]
        for pname, arr in params.items():
            cn = container_name(pname)
            if cn is None:
                unmapped.append(f"{part}: {pname}")
            elif cn.startswith(part + "."):
                tensors[cn] = _shape_fix(cn, arr)
    try:
        cfg = infer_config(tensors, name)
    except KeyError as e:
        raise ValueError(f"cannot infer the architecture, tensor missing: {e.args[0]}; unmapped: {unmapped[:20]}; unplaced: {unplaced[:20]}")
    want = weights.init_weights(cfg, 0)                 # the inventory the engine loads, with its shapes
    problems = [f"missing {k}" for k in want if k not in tensors]
    problems += [f"shape of {k}: {tensors[k].shape}, expected {want[k].shape}" for k in want if k in tensors and tensors[k].shape != want[k].shape]
    if problems:
        raise ValueError("ONNX files do not provide what the engine needs: " + "; ".join(problems[:40]) +
                         (f" ... (+{len(problems) - 40})" if len(problems) > 40 else ""))
    os.makedirs(dst_dir, exist_ok=True)
    paths = {}
    for part in ("encoder", "decoder", "joiner"):
        paths[part] = os.path.join(dst_dir, f"{part}-{name}.b200w")
        weights.save_container(paths[part], cfg, {k: tensors[k] for k in want}, prefix=part + ".")
    tok = os.path.join(src_dir, "tokens.txt")
    if os.path.exists(tok):
        paths["tokens"] = os.path.join(dst_dir, "tokens.txt")
        shutil.copyfile(tok, paths["tokens"])
    return {"config": cfg, "paths": paths, "unmapped": unmapped, "unplaced": unplaced}


if __name__ == "__main__":
    import sys
    res = convert_model_dir(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "zipformer-onnx")
    print("config:", res["config"])
    print("written:", res["paths"])
    print(f"{len(res['unmapped'])} named initialisers not used by the engine, {len(res['unplaced'])} anonymous matrices not placed")
