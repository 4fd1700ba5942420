"""Per-kernel roofline table for the C2 bench workload from an ncu launch list (one steady pass).

    python tools/kernel_rooflines.py gpurun_out/launches.csv > profiles/<name>.md

Algorithmic bytes / FLOPs come from the workload geometry, durations from ncu
(`gpu__time_duration.sum`, cold cache, serialised - an upper bound on the in-pipeline time), peaks from MEASURED_PEAKS.json.
"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from launch_summary import load  # noqa: E402
from sherpa_vietnamese_asr_b200 import synth, weights  # noqa: E402


def main():
    rows = load(sys.argv[1])
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tf = peaks["hbm_gbs"], peaks["bf16_tflops_sustained"]
    cfg = weights.CONFIGS["zipformer-68m"]()
    durs = synth.c2_durations(256, 256)
    n = [int(round(d * 16000)) for d in durs]
    T = [(x + 80) // 160 for x in n]
    T1 = [(t - 7) // 2 for t in T]
    t2 = [(t - 5) // 2 + 1 for t in T]
    Tk = {ds: [(t + ds - 1) // ds for t in T1] for ds in (1, 2, 4, 8)}
    M = {ds: sum(Tk[ds]) for ds in Tk}
    sq = {ds: sum(x * ((x + 3) & ~3) for x in Tk[ds]) for ds in Tk}
    M1 = M[1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    per = collections.defaultdict(list)
    for _, name, _, _, v in rows:
        agg[name][0] += 1
        agg[name][1] += v
        per[name].append(v)
    stacks = list(zip(cfg.num_encoder_layers, cfg.downsampling_factor, cfg.encoder_dim, cfg.feedforward_dim, cfg.num_heads,
                      cfg.cnn_module_kernel))
    # algorithmic bytes per kernel family (whole pass)
    b = {}
    b["fbank_kernel"] = 4 * sum(n) + 320 * sum(T)
    b["embed_conv0_kernel"] = 4 * (80 * sum(T) + 640 * sum(t - 2 for t in T))
    b["embed_conv1_kernel"] = 4 * (640 * sum(t - 2 for t in T) + 39 * 32 * sum(t2))
    b["embed_im2col2_kernel"] = 4 * (39 * 32 * sum(t2) + 19 * 288 * M1)
    b["embed_dw7_kernel"] = 4 * 2 * 19 * 128 * M1
    b["glu_dwconv"] = sum(L * 2 * 4 * M[ds] * 3 * D for L, ds, D, F, H, k in stacks)
    b["attn_weights_tcgen05_kernel"] = sum(L * 4 * (H * sq[ds] + M[ds] * H * 68) for L, ds, D, F, H, k in stacks)
    b["attn_apply_tcgen05_kernel<16"] = sum(L * 2 * 4 * (H * sq[ds] + 2 * M[ds] * H * 12) for L, ds, D, F, H, k in stacks)
    b["attn_apply_tcgen05_kernel<64"] = sum(L * 4 * (sq[ds] + 4 * M[ds] * (3 * D // 4)) for L, ds, D, F, H, k in stacks)
    b["upsample_combine_kernel"] = sum(4 * (2 * M1 + M[ds]) * D for L, ds, D, F, H, k in stacks if ds > 1)
    b["biasnorm_kernel"] = sum(L * 4 * 3 * M[ds] * D for L, ds, D, F, H, k in stacks) + 4 * 2 * M1 * cfg.encoder_dim[0]
    b["bypass_kernel"] = sum(L * 4 * 3 * M[ds] * D for L, ds, D, F, H, k in stacks)
    fl_gemm = 0.0
    fl_gemm += 2.0 * M1 * 19 * (128 * 288 + 384 * 128 + 128 * 384) + 2.0 * M1 * cfg.encoder_dim[0] * 2432
    for L, ds, D, F, H, k in stacks:
        h = 3 * D // 4
        per_row = (D * 68 * H + 2 * D * (3 * F // 4 + F + 5 * F // 4) + D * 3 * h + h * D + 2 * (2 * D * 12 * H) + 2 * (2 * D * D + D * D))
        fl_gemm += 2.0 * L * M[ds] * per_row
    fl_gemm += 2.0 * M[2] * max(cfg.encoder_dim) * cfg.joiner_dim
    fl_aw = sum(L * 2.0 * 36 * H * sum(x * x for x in Tk[ds]) for L, ds, D, F, H, k in stacks)
    print("# Per-kernel rooflines, C2 workload (256 segments, 2856 audio-s, Zipformer-68M, beam 4, FP32 mode)\n")
    print(f"Launch list: `{os.path.basename(sys.argv[1])}`; peaks of measured (MEASURED_PEAKS.json): HBM {hbm:.0f} GB/s, dense BF16 "
          f"{tf:.0f} TFLOP/s sustained. Durations are ncu `gpu__time_duration.sum` (cold cache, serialised), so the fractions are "
          "lower bounds of what the kernels reach inside the pipeline.\n")
    print("| kernel family | launches | total ms | bound | algorithmic work | achieved | fraction of peak |")
    print("|---|---:|---:|---|---|---|---:|")

    def fam(prefix):
        c = sum(v[0] for k, v in agg.items() if k.startswith(prefix))
        t = sum(v[1] for k, v in agg.items() if k.startswith(prefix))
        return c, t

    for key, by in b.items():
        c, t = fam(key)
        if not c:
            continue
        gbs = by / 1e9 / (t * 1e-6)
        print(f"| `{key}` | {c} | {t / 1e3:.2f} | HBM | {by / 1e9:.2f} GB | {gbs:.0f} GB/s | {gbs / hbm:.3f} |")
    c = t = 0
    for pre in ("gemm_tf32_tcgen05_kernel<128, 1, 0", "gemm_tf32_tcgen05_kernel<64, 1, 0", "gemm_tf32_tcgen05_kernel<128, 1, -",
                "gemm_tf32_tcgen05_kernel<64, 1, -"):   # EPI 0 / -1 / -2 = encoder Linears (none / SwooshL / SwooshR)
        cc, tt = fam(pre)
        c, t = c + cc, t + tt
    if c:
        a = fl_gemm / 1e12 / (t * 1e-6)
        print(f"| `gemm_tf32_tcgen05_kernel` (encoder Linears, 3xTF32) | {c} | {t / 1e3:.2f} | tensor | {fl_gemm / 1e12:.2f} TFLOP (x3 MMAs issued) | "
              f"{a:.0f} TFLOP/s | {a / tf:.3f} (x3 = {3 * a / tf:.3f} of the BF16 peak in issued MMA FLOPs) |")
    c = t = 0
    for pre in ("gemm_f16_tcgen05_kernel<128, 0", "gemm_f16_tcgen05_kernel<64, 0", "gemm_f16_tcgen05_kernel<128, -", "gemm_f16_tcgen05_kernel<64, -"):
        cc, tt = fam(pre)
        c, t = c + cc, t + tt
    if c:
        a = fl_gemm / 1e12 / (t * 1e-6)
        print(f"| `gemm_f16_tcgen05_kernel` (encoder Linears, fp16 hi/lo operand split) | {c} | {t / 1e3:.2f} | tensor | {fl_gemm / 1e12:.2f} TFLOP (x3 MMAs issued) | "
              f"{a:.0f} TFLOP/s | {a / tf:.3f} (x3 = {3 * a / tf:.3f} of the BF16 peak in issued MMA FLOPs) |")
    c, t = fam("attn_weights_tcgen05_kernel")
    if c:
        print(f"| `attn_weights_tcgen05_kernel` (as FLOPs) | {c} | {t / 1e3:.2f} | tensor | {fl_aw / 1e12:.3f} TFLOP | {fl_aw / 1e12 / (t * 1e-6):.1f} TFLOP/s | "
              f"{fl_aw / 1e12 / (t * 1e-6) / tf:.4f} (epilogue-bound: K = 32) |")
    js, jt = fam("joiner_f16ss_tcgen05_kernel")
    if js:
        st = fam("select_partials_kernel")[1]
        print(f"| search step: `joiner_f16ss` GEMM / `select_partials` (decoder table: no decoder kernel) | {js} steps | {(jt + st) / 1e3:.2f} | latency | "
              f"{jt / js:.1f} / {st / js:.1f} us per frame step (cold cache; 21 us per step live) | - | - |")
    steps = agg.get("decoder_joinin_kernel", [0, 0])[0]
    if steps:
        tj = fam("gemm_tf32_tcgen05_kernel<64, 1, 4")[1] + fam("gemm_tf32_tcgen05_kernel<128, 1, 4")[1]
        print(f"| search step: `decoder_joinin` / joiner GEMM / `select_partials` | {steps} steps | "
              f"{(agg['decoder_joinin_kernel'][1] + tj + fam('select_partials_kernel')[1]) / 1e3:.2f} | latency | "
              f"{agg['decoder_joinin_kernel'][1] / steps:.1f} / {tj / steps:.1f} / {fam('select_partials_kernel')[1] / steps:.1f} us per frame step | - | - |")


if __name__ == "__main__":
    main()
