"""Times the window-attention op alone on the four Swin-T stage shapes of the benchmark batch (192 images), plain and
shifted:  python tools/winattn_time.py [N] [bf16].  With a library built by
`tools/build_variant.sh probes svx_winattn -DSVX_WINATTN_PROBES` and SVX_WINATTN_PROBE=<bits> (1 no loads, 2 no softmax,
8 no P.V, 16 no stores, 32 no V conversion, 128 strictly ordered MMA issue) it shows where the item time goes -- the
results of a probed run are WRONG by construction (profiles/r2_winattn_probes.txt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from swinvox_b200 import engine as E  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 192
bf = len(sys.argv) > 2
dt = torch.bfloat16 if bf else torch.float32
torch.manual_seed(1)
out_line = []
for H, heads in ((56, 3), (28, 6), (14, 12), (7, 24)):
    C = heads * 32
    for shift in ((0, 3) if H > 7 else (0,)):
        p = E.Plan("cuda", dtype=dt)
        qkv = p.new_act(N, 1, H, H, 3 * C)
        qkv.buf.copy_(torch.randn(qkv.buf.shape, device="cuda"))
        out = p.new_act(N, 1, H, H, C)
        p.window_attention(qkv, out, (torch.randn(heads, 49, 49) * 0.5).cuda(), H, H, heads, shift, 32 ** -0.5)
        for _ in range(3):
            p.run()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(10):
            p.run()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 100
        gbs = 4 * qkv.esize * N * H * H * C / us / 1e3
        out_line.append(f"H{H}{'s' if shift else ''} {us:.0f} us ({gbs:.0f} GB/s)")
print(f"probe {os.environ.get('SVX_WINATTN_PROBE', '0'):>3s} {'bf16' if bf else 'fp32'} N={N}: " + "  ".join(out_line))
