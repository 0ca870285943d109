"""Runs one window-attention configuration on the GPU against the fp64 reference (separate process per case, so that a
launch failure is attributed to its case):  python tools/winattn_cases.py H heads shift N [bf16]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from swinvox_b200 import engine as E  # noqa: E402

H, heads, shift, N = [int(v) for v in sys.argv[1:5]]
bf = len(sys.argv) > 5
C = heads * 32
torch.manual_seed(1)
qkv = torch.randn(N, H, H, 3 * C)
qkv = qkv.to(torch.bfloat16).float() if bf else E.tf32_round(qkv)
bias = torch.randn(heads, 49, 49) * 0.5
p = E.Plan("cuda", dtype=torch.bfloat16 if bf else torch.float32)
out = p.new_act(N, 1, H, H, C)
src = qkv.reshape(-1, 3 * C)
p.window_attention(E.Act((src.to(torch.bfloat16) if bf else src).cuda(), N, 1, H, H, 3 * C), out, bias.cuda(), H, H, heads, shift,
                   32 ** -0.5, round_out=False)
for _ in range(3):
    p.run()
torch.cuda.synchronize()
x = torch.roll(qkv.double(), (-shift, -shift), (1, 2)) if shift else qkv.double()
nw = H // 7
xw = x.view(N, nw, 7, nw, 7, 3, heads, 32).permute(0, 1, 3, 5, 6, 2, 4, 7).reshape(N * nw * nw, 3, heads, 49, 32)
q, k, v = xw[:, 0], xw[:, 1], xw[:, 2]
att = (q * 32 ** -0.5) @ k.transpose(-1, -2) + bias.double()
if shift:
    m = torch.zeros(H, H)
    for i, hs in enumerate((slice(0, -7), slice(-7, -shift), slice(-shift, None))):
        for j, ws_ in enumerate((slice(0, -7), slice(-7, -shift), slice(-shift, None))):
            m[hs, ws_] = i * 3 + j
    mw = m.view(nw, 7, nw, 7).permute(0, 2, 1, 3).reshape(nw * nw, 49)
    att = att + ((mw[:, None, :] != mw[:, :, None]).double() * -100.0).repeat(N, 1, 1)[:, None]
o = (att.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(N, nw, nw, 7, 7, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H, H, C)
if shift:
    o = torch.roll(o, (shift, shift), (1, 2))
err = ((out.view().squeeze(1).float().cpu().double() - o).abs().max() / o.abs().max()).item()
print(f"H={H} heads={heads} shift={shift} N={N} {'bf16' if bf else 'fp32'}: rel err {err:.2e}")
