"""runs one implicit-GEMM convolution a few times (for ncu): python tools/one_conv.py N C H Cout k stride bn"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from swinvox_b200 import engine as E  # noqa: E402

N, C, H, Cout, k, s, bn = [int(v) for v in sys.argv[1:8]]
x = E.tf32_round(torch.randn(N * H * H, C, device="cuda"))
conv = torch.nn.Conv2d(C, Cout, k, s, k // 2).cuda()
p = E.Plan("cuda")
oh = (H + 2 * (k // 2) - k) // s + 1
out = p.new_act(N, 1, oh, oh, Cout)
p.conv(E.Act(x, N, 1, H, H, C), E.pack_conv(conv.weight, conv.bias, None, "cuda", block_n=bn),
       E.conv_taps(1, k, k, 0, k // 2, k // 2), out, stride=(1, s, s), act=E.ACT_RELU)
for _ in range(3):
    p.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(5):
    p.run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"conv N={N} C={C} H={H} Cout={Cout} k={k} s={s} bn={bn}: {ms:.4f} ms  {2.0 * N * oh * oh * Cout * C * k * k / ms / 1e9:.1f} TF/s")
