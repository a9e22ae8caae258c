"""Per-GEMM roofline table for one step at C* (B = 64, fp32 storage / TF32): pairs the launch order of the forward
schedule with an ncu launch list.  ideal = max(FLOPs / tf32_peak, unique HBM bytes / hbm_bw).
Usage: python profiles/gemm_roofline.py launches.csv"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize_launches import load, short  # noqa: E402

B, ES = 64, 4
PK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else \
    {"bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
TF32 = PK["bf16_tflops_sustained"] / 2 * 1e12
TF32_BURST = PK.get("bf16_tflops", PK["bf16_tflops_sustained"]) / 2 * 1e12   # the peak at the clock a short run holds
HBM = PK["hbm_gbs"] * 1e9
FOLD = os.environ.get("FOLD", "1") == "1"       # round 2: algebraically folded neck (SPE_FOLD_NECK, default on)
KV_X3 = os.environ.get("KV_X3", "0") == "1"     # round 2: the K/V projection is plain TF32 once calibrated
FUSE_DOWN = os.environ.get("FUSE_DOWN", "1") == "1"   # round 2 (r02d on): conv3 + downsample of a layer's first block as one GEMM
DEC0_FOLD = os.environ.get("DEC0_FOLD", "1") == "1"   # round 2 (r02c on): decoder layer 0's query-only projections precomputed


def schedule():
    g = []  # (name, M, N, K, extra_read_elems)

    def add(name, M, N, K, res=0, x3=False):
        g.append((name, M, N, K, res, x3, 1))
    add("stem 7x7/s2", B * 112 * 112, 64, 147)
    inpl, H = 64, 56
    for li, (pl, nb) in enumerate(((64, 3), (128, 4), (256, 6))):
        for bi in range(nb):
            s = 2 if (bi == 0 and li > 0) else 1
            Ho = H // s
            add(f"l{li + 1}.{bi}.conv1", B * H * H, pl, inpl)
            add(f"l{li + 1}.{bi}.conv2 3x3", B * Ho * Ho, pl, 9 * pl)
            if bi == 0 and FUSE_DOWN:
                # out = relu([t | x sampled at the stride] . [s3 W3 | s_d W_d]^T + b): K = planes + inplanes
                add(f"l{li + 1}.{bi}.conv3|down (K-concat)", B * Ho * Ho, 4 * pl, pl + inpl)
                inpl, H = 4 * pl, Ho
                continue
            if bi == 0:
                add(f"l{li + 1}.{bi}.down", B * Ho * Ho, 4 * pl, inpl)
            add(f"l{li + 1}.{bi}.conv3+res", B * Ho * Ho, 4 * pl, pl, res=B * Ho * Ho * 4 * pl)
            inpl, H = 4 * pl, Ho
    T = 784
    add("s8_latern", B * T, 256, 512)
    if FOLD:
        # round 2: s16_latern's nine tap matrices on the 14 x 14 map (then upsample_tapsum_kernel); output_conv and
        # input_proj as ONE 3x3 convolution 512 -> 256
        add("s16_latern taps @14x14", B * 196, 9 * 256, 1024)
        add("output_conv.input_proj 3x3", B * T, 256, 9 * 512)
    else:
        add("s16_latern 3x3", B * T, 256, 9 * 1024)
        add("output_conv 3x3", B * T, 512, 9 * 512)
        add("input_proj", B * T, 256, 512)
    for i in range(4):
        add(f"enc{i}.qkv", B * T, 768, 256)
        add(f"enc{i}.out+res", B * T, 256, 256, res=B * T * 256)
        # linear1 + ReLU + linear2 + residual + norm2 in one kernel: 2 x (M x 2048 x 256) MACs, X in, Y out
        g.append((f"enc{i}.ffn fused (+LN)", B * T, 2048, 256, 0, False, 2))
    add("dec.kv_all" + (" (3xTF32)" if KV_X3 else ""), B * T, 2048, 256, x3=KV_X3)
    Q = 40
    for i in range(4):
        if not (i == 0 and DEC0_FOLD):
            add(f"dec{i}.sa_qkv x3", B * Q, 768, 256, x3=True)
            add(f"dec{i}.sa_out x3", B * Q, 256, 256, x3=True)
            add(f"dec{i}.ca_q x3", B * Q, 256, 256, x3=True)
        add(f"dec{i}.ca_out x3", B * Q, 256, 256, x3=True)
        add(f"dec{i}.ff1 x3", B * Q, 2048, 256, x3=True)
        add(f"dec{i}.ff2 x3", B * Q, 256, 2048, x3=True)
    add("pt0 x3", B * Q, 256, 256, x3=True)
    add("pt1 x3", B * Q, 256, 256, x3=True)
    return g


def main(path):
    extra = {}
    seq = load(path, extra)
    starts = [i for i, s in enumerate(seq) if "crop_resize" in s[1]]
    ends = [i for i, s in enumerate(seq) if "assign_pnp" in s[1] or "PnpDesc" in s[1]]
    a = starts[-1] if ends and ends[-1] > starts[-1] else starts[-2]
    b = min(e for e in ends if e > a)
    FAM = ("gemm_tc_kernel", "gemm_tc2_kernel", "conv3_tc_kernel", "ffn_tc2_kernel", "ffn_tc_kernel")
    gem = [s for s in seq[a:b + 1] if short(s[1]) in FAM]
    sch = schedule()
    assert len(gem) == len(sch), (len(gem), len(sch))
    TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
    has_tp = any(TP in v for v in extra.values())
    print("| GEMM | M | N | K | us | ideal us | TFLOP/s | x ideal |" + (" tensor pipe % of elapsed |" if has_tp else "") +
          "\n|---|---:|---:|---:|---:|---:|---:|---:|" + ("---:|" if has_tp else ""))
    tot = tot_ideal = tot_burst = 0
    tp_weighted = 0.0
    tag = {"gemm_tc2_kernel": " (pair)", "conv3_tc_kernel": " (tap reuse)", "ffn_tc_kernel": "", "ffn_tc2_kernel": " (pair)"}
    for (name, M, N, K, res, x3, nmm), s in zip(sch, gem):
        fl = 2 * M * N * K * nmm
        by = (M * K + M * N + res) * ES if "3x3" not in name else (M * K // 9 + M * N) * ES
        if nmm == 2:
            by = 2 * M * K * ES
        name = name + tag.get(short(s[1]), "")
        ideal = max(fl * (3 if x3 else 1) / TF32, by / HBM) * 1e6
        tot += s[2]; tot_ideal += ideal
        tot_burst += max(fl * (3 if x3 else 1) / TF32_BURST, by / HBM) * 1e6
        tp = extra.get(s[0], {}).get(TP)
        if tp is not None:
            tp_weighted += tp * s[2]
        print(f"| {name} | {M} | {N} | {K} | {s[2]:.1f} | {ideal:.1f} | {fl / s[2] / 1e6:.0f} | {s[2] / ideal:.1f} |" +
              (f" {tp:.1f} |" if tp is not None else ""))
    print(f"\nGEMM total {tot:.0f} us, sum of per-GEMM ideals {tot_ideal:.0f} us at the sustained peak "
          f"({TF32 / 1e12:.0f} TFLOP/s), {tot_burst:.0f} us at the burst peak ({TF32_BURST / 1e12:.0f} TFLOP/s)")
    if has_tp:
        print(f"time-weighted tensor-pipe utilisation of the GEMM family (counter {TP}): {tp_weighted / tot:.1f} %")
    # the other kernels that issue tensor-core work
    att = [s for s in seq[a:b + 1] if "attention" in s[1]]
    if has_tp and att:
        t = sum(s[2] for s in att)
        w = sum(extra.get(s[0], {}).get(TP, 0.0) * s[2] for s in att)
        step_t = sum(s[2] for s in seq[a:b + 1])
        allw = sum(extra.get(s[0], {}).get(TP, 0.0) * s[2] for s in seq[a:b + 1])
        print(f"attention kernels: {t:.0f} us at {w / t:.1f} %; whole step ({step_t:.0f} us of kernel time): {allw / step_t:.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
