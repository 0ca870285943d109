"""Times the tcgen05 contraction kernel on the SwinVox problem shapes (cfg2: 192 images) and prints
achieved TFLOP/s next to torch.matmul (cuBLAS TF32) on the same shape.  Run on the GPU box."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swinvox_b200 import engine as E

DEV = "cuda"


def time_cuda(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def bench_linear(M, N, K, bn):
    x = E.tf32_round(torch.randn(M, K, device=DEV))
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    p = E.Plan(DEV)
    out = p.new_act(M, 1, 1, 1, N)
    p.linear(E.Act(x, M, 1, 1, 1, K), E.pack_matrix(w, None, DEV, block_n=bn), out)
    ms = time_cuda(p.run)
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_t = time_cuda(lambda: torch.matmul(x, w.t()))
    fl = 2.0 * M * N * K
    return ms, fl / ms / 1e9, ms_t, fl / ms_t / 1e9


def bench_conv(N, C, H, Cout, k, s, bn):
    x = E.tf32_round(torch.randn(N * H * H, C, device=DEV))
    conv = torch.nn.Conv2d(C, Cout, k, s, k // 2).to(DEV)
    p = E.Plan(DEV)
    oh = (H + 2 * (k // 2) - k) // s + 1
    out = p.new_act(N, 1, oh, oh, Cout)
    p.conv(E.Act(x, N, 1, H, H, C), E.pack_conv(conv.weight, conv.bias, None, DEV, block_n=bn),
           E.conv_taps(1, k, k, 0, k // 2, k // 2), out, stride=(1, s, s), act=E.ACT_RELU)
    ms = time_cuda(p.run)
    xt = x.view(N, H, H, C).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    convc = conv.to(memory_format=torch.channels_last)
    torch.backends.cudnn.allow_tf32 = True
    with torch.no_grad():
        ms_t = time_cuda(lambda: convc(xt))
    fl = 2.0 * N * oh * oh * Cout * C * k * k
    return ms, fl / ms / 1e9, ms_t, fl / ms_t / 1e9


if __name__ == "__main__":
    torch.manual_seed(0)
    # TF32 peak of this GPU by the MEASURED_PEAKS.json protocol (torch.matmul 8192^3, best of 10)
    a = torch.randn(8192, 8192, device=DEV)
    b = torch.randn(8192, 8192, device=DEV)
    torch.backends.cuda.matmul.allow_tf32 = True
    best = min(time_cuda(lambda: torch.matmul(a, b), iters=1, warm=1) for _ in range(10))
    print(json.dumps({"tf32_tflops_burst_cublas_8192": 2 * 8192 ** 3 / best / 1e9}))
    del a, b
    rows = []
    for (M, N, K) in [(602112, 288, 96), (602112, 96, 96), (602112, 384, 96), (602112, 96, 384),
                      (150528, 576, 192), (150528, 768, 192), (150528, 192, 768),
                      (37632, 1152, 384), (37632, 1536, 384), (37632, 384, 1536), (37632, 256, 1024),
                      (9408, 2304, 768), (9408, 3072, 768), (9408, 768, 3072),
                      (602112, 64, 256), (602112, 256, 64), (602112, 64, 64), (150528, 512, 128), (150528, 128, 512), (150528, 128, 256), (37632, 1024, 256), (37632, 256, 1024), (9408, 256, 768), (64, 2048, 8192), (64, 8192, 2048),
                      (8192, 8192, 8192)]:
        for bn in ([64] if N == 64 else [64, 96, 128, 192, 256]):
            if N % bn and not (N < bn):
                continue
            if N < bn and bn != 96 and bn != 128:
                continue
            try:
                ms, tf, ms_t, tf_t = bench_linear(M, N, K, bn)
                rows.append(dict(kind="linear", M=M, N=N, K=K, bn=bn, ms=round(ms, 4), tflops=round(tf, 1),
                                 torch_ms=round(ms_t, 4), torch_tflops=round(tf_t, 1)))
                print(json.dumps(rows[-1]), flush=True)
            except Exception as e:  # noqa
                print("FAILED", M, N, K, bn, e, flush=True)
    for (N, C, H, Cout, k, s) in [(192, 64, 56, 64, 3, 1), (192, 128, 56, 128, 3, 2), (192, 256, 14, 256, 3, 1),
                                  (192, 256, 56, 256, 3, 2), (192, 512, 7, 256, 3, 1), (192, 256, 7, 256, 3, 1)]:
        for bn in [64, 128, 256]:
            if Cout % bn:
                continue
            try:
                ms, tf, ms_t, tf_t = bench_conv(N, C, H, Cout, k, s, bn)
                rows.append(dict(kind="conv", N=N, C=C, H=H, Cout=Cout, k=k, s=s, bn=bn, ms=round(ms, 4),
                                 tflops=round(tf, 1), torch_ms=round(ms_t, 4), torch_tflops=round(tf_t, 1)))
                print(json.dumps(rows[-1]), flush=True)
            except Exception as e:  # noqa
                print("FAILED conv", N, C, H, Cout, k, s, bn, e, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)
