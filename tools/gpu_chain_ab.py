"""A/B timing of the persistent chain kernels (tf32 vs fp16 formulation) at the three cfg3 stage shapes.  Not a pytest.
usage: python tools/gpu_chain_ab.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    for prec in ("fast_tf32", "fast_f16"):
        for r in bench.kernel_microbench(torch, prec, B):
            print("%-9s %-40s %-20s %8.1f us  %7.1f GB/s  %6.1f TFLOP/s" % (prec, r["kernel"], r["shape"], r["us"], r["GBps"], r["algorithmic_TFLOPs"]), flush=True)
