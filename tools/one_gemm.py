"""runs one contraction a few times (for ncu): python tools/one_gemm.py M N K bn [act] [res]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swinvox_b200 import engine as E
M, N, K, bn = [int(v) for v in sys.argv[1:5]]
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
res = int(sys.argv[6]) if len(sys.argv) > 6 else 0
x = E.tf32_round(torch.randn(M, K, device="cuda"))
w = torch.randn(N, K, device="cuda") / K ** 0.5
p = E.Plan("cuda")
out = p.new_act(M, 1, 1, 1, N)
r = E.Act(torch.randn(M, N, device="cuda"), M, 1, 1, 1, N) if res else None
p.linear(E.Act(x, M, 1, 1, 1, K), E.pack_matrix(w, torch.randn(N), "cuda", block_n=bn), out, act=act, residual=r)
for _ in range(3):
    p.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(5):
    p.run()
b.record()
torch.cuda.synchronize()
print(f"M={M} N={N} K={K} bn={bn} act={act} res={res}: {a.elapsed_time(b)/5:.4f} ms")
