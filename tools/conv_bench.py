import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.gemm_bench import bench_conv
for (N, C, H, Cout, k, s, bn) in [(192, 64, 56, 64, 3, 1, 64), (192, 256, 14, 256, 3, 1, 256), (192, 256, 14, 256, 3, 1, 128), (192, 512, 7, 256, 3, 1, 128)]:
    ms, tf, ms_t, tf_t = bench_conv(N, C, H, Cout, k, s, bn)
    print(f"conv C={C} H={H} Co={Cout} bn={bn}: {ms:.4f} ms {tf:.1f} TF | torch {ms_t:.4f} ms", flush=True)
