"""ONNX initialisers -> .b200w containers (sherpa-vietnamese-asr_b200/onnx_import.py). Neither `onnx` nor a real checkpoint exists
offline, so this is a SELF-CONSISTENCY test: the seeded tiny model is written as three ONNX files in the naming icefall's
export is known to use - module-path names for convolutions / biases / norm parameters, Linear weights as anonymous transposed
`onnx::MatMul_<n>` initialisers reachable through their MatMul node (by node name, or through the bias of the consuming Add) -
and must come back tensor for tensor, with the architecture inferred from the shapes alone. It pins the wire-format reader and
the name mapping, not the claim that real exports look like this."""
import os
import re

import numpy as np
import pytest

from sherpa_vietnamese_asr_b200 import onnx_import as oi
from sherpa_vietnamese_asr_b200 import weights

_LAYER_BACK = {"attn_w.in_proj": "self_attn_weights.in_proj", "attn_w.linear_pos": "self_attn_weights.linear_pos",
               "nonlin.in": "nonlin_attention.in_proj", "nonlin.out": "nonlin_attention.out_proj", "norm": "norm"}
_EMBED_BACK = {v: k for k, v in oi._EMBED.items()}


def _icefall_path(cn: str) -> str:
    part, rest = cn.split(".", 1)
    if part == "decoder":
        return {"embedding": "decoder.embedding", "conv": "decoder.conv", "decoder_proj": "decoder_proj"}[rest.rsplit(".", 1)[0]] + "." + rest.rsplit(".", 1)[1]
    if part == "joiner":
        return rest
    if rest.startswith("embed."):
        mod, param = rest[len("embed."):].rsplit(".", 1)
        return f"encoder_embed.{_EMBED_BACK[mod]}.{param}"
    if rest.startswith("encoder_proj."):
        return rest
    if rest == "downsample_output.bias":
        return "encoder.downsample_output.bias"
    m = re.fullmatch(r"stack(\d+)\.(.+)", rest)
    i, tail = int(m.group(1)), m.group(2)
    if tail == "downsample.bias":
        return f"encoder.encoders.{i}.downsample.bias"
    if tail == "out_combiner.scale":
        return f"encoder.encoders.{i}.out_combiner.bypass_scale"
    m = re.fullmatch(r"layer(\d+)\.(.+)\.(\w+)", tail)
    l, mod, param = m.group(1), m.group(2), m.group(3)
    if mod in ("bypass", "bypass_mid"):
        mod, param = mod, "bypass_scale"
    elif mod in _LAYER_BACK:
        mod = _LAYER_BACK[mod]
    else:
        mm = re.fullmatch(r"(ff|attn|conv)([123])\.(in|out|dw)", mod)
        kind = {"ff": "feed_forward", "attn": "self_attn", "conv": "conv_module"}[mm.group(1)]
        mod = f"{kind}{mm.group(2)}." + {"in": "in_proj", "out": "out_proj", "dw": "depthwise_conv"}[mm.group(3)]
    return f"encoder.encoders.{i}." + ("encoder." if i > 0 else "") + f"layers.{l}.{mod}.{param}"


def _write_part(path, W, part):
    inits, nodes, k = {}, [], 0
    for cn, a in W.items():
        if not cn.startswith(part + "."):
            continue
        ip = _icefall_path(cn)
        a = np.asarray(a, dtype=np.float32)
        is_linear = ip.endswith(".weight") and a.ndim == 2 and "embedding" not in ip
        if cn.endswith(("convnext.pw1.weight", "convnext.pw2.weight")):
            inits[ip] = a.reshape(a.shape[0], a.shape[1], 1, 1)               # 1x1 Conv2d in the exported graph
        elif is_linear:
            anon = f"onnx::MatMul_{1000 + k}"
            k += 1
            inits[anon] = np.ascontiguousarray(a.T)
            module = ip[: -len(".weight")]
            has_bias = (cn[: -len("weight")] + "bias") in W
            if has_bias and part != "encoder":     # Linear on 2-D inputs: Gemm(A, B [out, in] or [in, out], C = named bias)
                if part == "decoder":
                    inits[anon] = np.ascontiguousarray(a)                      # transB = 1 form
                nodes.append({"op_type": "Gemm", "name": f"Gemm_{k}", "inputs": [f"x{k}", anon, module + ".bias"], "outputs": [f"y{k}"]})
            elif has_bias and k % 2:        # half of the biased Linears: anonymous node name, found through the Add's bias
                nodes.append({"op_type": "MatMul", "name": f"MatMul_{k}", "inputs": [f"x{k}", anon], "outputs": [f"mm{k}"]})
                nodes.append({"op_type": "Add", "name": f"Add_{k}", "inputs": [module + ".bias", f"mm{k}"], "outputs": [f"y{k}"]})
            else:
                nodes.append({"op_type": "MatMul", "name": re.sub(r"/(\d+)(?=/|$)", r".\1", "/" + module.replace(".", "/")) + "/MatMul",
                              "inputs": [f"x{k}", anon], "outputs": [f"mm{k}"]})
        else:
            inits[ip] = a.reshape(()) if cn.endswith("log_scale") else a
    inits["unrelated.running_mean"] = np.zeros(3, np.float32)
    inits["onnx::Constant_7"] = np.array([2, 3], np.int64)
    oi.write_model(path, inits, nodes)


def test_module_paths_map_to_container_names():
    assert oi.container_name("encoder.encoders.0.layers.1.feed_forward3.out_proj.bias") == "encoder.stack0.layer1.ff3.out.bias"
    assert oi.container_name("encoder.encoders.3.encoder.layers.2.conv_module2.depthwise_conv.weight") == "encoder.stack3.layer2.conv2.dw.weight"
    assert oi.container_name("encoder.encoders.2.encoder.layers.0.bypass_mid.bypass_scale") == "encoder.stack2.layer0.bypass_mid.scale"
    assert oi.container_name("encoder.encoders.4.out_combiner.bypass_scale") == "encoder.stack4.out_combiner.scale"
    assert oi.container_name("encoder_embed.conv.7.weight") == "encoder.embed.conv2.weight"
    assert oi.container_name("encoder_embed.out_norm.log_scale") == "encoder.embed.out_norm.log_scale"
    assert oi.container_name("decoder_proj.weight") == "decoder.decoder_proj.weight"
    assert oi.container_name("encoder.encoders.0.layers.0.balancer1.min_positive") is None
    assert oi._module_of_node("/encoder/encoders.1/encoder/layers.0/self_attn1/in_proj/MatMul") == "encoder.encoders.1.encoder.layers.0.self_attn1.in_proj"


def test_synthetic_onnx_round_trip(tmp_path):
    cfg = weights.zipformer_tiny()
    W = weights.init_weights(cfg, 11)
    src, dst = tmp_path / "onnx", tmp_path / "b200w"
    os.makedirs(src)
    for part in ("encoder", "decoder", "joiner"):
        _write_part(str(src / f"{part}-epoch-1-avg-1.onnx"), W, part)
    with open(src / "tokens.txt", "w", encoding="utf-8") as f:
        f.write("<blk> 0\n")
    res = oi.convert_model_dir(str(src), str(dst), "tiny-from-onnx")
    got_cfg = res["config"]
    for field in ("num_encoder_layers", "downsampling_factor", "encoder_dim", "feedforward_dim", "num_heads", "cnn_module_kernel",
                  "query_head_dim", "pos_head_dim", "value_head_dim", "pos_dim", "decoder_dim", "joiner_dim", "context_size", "vocab_size"):
        assert getattr(got_cfg, field) == getattr(cfg, field), field
    back = {}
    for part in ("encoder", "decoder", "joiner"):
        back.update(weights.load_container(res["paths"][part])[1])
    assert set(back) == set(W)
    for k in W:
        np.testing.assert_array_equal(back[k], W[k], err_msg=k)
    assert os.path.exists(res["paths"]["tokens"])
    assert any("running_mean" in u for u in res["unmapped"]) and not res["unplaced"]


def test_missing_tensors_are_reported_not_guessed(tmp_path):
    cfg = weights.zipformer_tiny()
    W = weights.init_weights(cfg, 12)
    del W["encoder.stack2.layer0.norm.log_scale"]            # e.g. folded into a constant by the exporter
    del W["encoder.stack1.layer0.ff2.out.bias"]
    src = tmp_path / "onnx"
    os.makedirs(src)
    for part in ("encoder", "decoder", "joiner"):
        _write_part(str(src / f"{part}-x.onnx"), W, part)
    with pytest.raises(ValueError) as e:
        oi.convert_model_dir(str(src), str(tmp_path / "out"))
    assert "encoder.stack2.layer0.norm.log_scale" in str(e.value) and "encoder.stack1.layer0.ff2.out.bias" in str(e.value)
