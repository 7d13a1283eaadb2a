"""Markdown table of a cfg2 sweep (tools/gpu_sweep.py --out ...jsonl) against the roof that binds each row
(SURVEY.md section 8d): HBM (measured copy bandwidth) for C <= 32 in the fast modes (C <= 64 with fp32 I/O), the
tensor roof above -- measured cuBLAS bf16 burst peak for fast_bf16, half of it for tf32 operands (no measured tf32
peak exists: labelled "est."), a third of that for strict (3xTF32 issues three MMAs per algorithmic one).
usage: python tools/sweep_table.py profiles/r01_sweep_cfg2_v6.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM, BF16 = peaks["hbm_gbs"], peaks["bf16_tflops"]
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print("| C | mode | kernel | µs | alg. GB/s | alg. TFLOP/s | binding roof | fraction |")
print("|---|---|---|---|---|---|---|---|")
for r in sorted(rows, key=lambda r: (r["shape"][3], ["fast_bf16", "fast_tf32", "strict"].index(r["mode"]), ["fwd", "dgrad", "wgrad"].index(r["kernel"]))):
    C, mode = r["shape"][3], r["mode"]
    if "us" not in r:
        continue
    hbm_bound = C <= 32 if mode == "fast_bf16" else (C <= 64 if mode == "fast_tf32" else C <= 16)
    if hbm_bound:
        roof, frac = "HBM %.0f GB/s (measured)" % HBM, r["GBps"] / HBM
    else:
        peak = BF16 if mode == "fast_bf16" else BF16 / 2 if mode == "fast_tf32" else BF16 / 6
        label = "measured" if mode == "fast_bf16" else "est. bf16/2" if mode == "fast_tf32" else "est. bf16/6"
        roof, frac = "tensor %.0f TFLOP/s (%s)" % (peak, label), r["alg_TFLOPs"] / peak
    print("| %d | %s | %s | %.1f | %.0f | %.1f | %s | %.0f %% |" % (C, mode, r["kernel"], r["us"], r["GBps"], r["alg_TFLOPs"], roof, 100 * frac))
