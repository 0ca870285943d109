"""times the fused MLP kernel (svx_mlp.cu) on the Swin stage-0 / stage-1 shapes of the 64 x 3 view workload.
usage: SVX_LIB_PATH=<variant .so> python tools/probes/mlp_time.py [label]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from swinvox_b200 import engine as E  # noqa: E402

dev = torch.device("cuda:0")
label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("SVX_LIB_PATH", "default")
for M, Cc in ((192 * 56 * 56, 96), (192 * 28 * 28, 192)):
    torch.manual_seed(0)
    hid = 4 * Cc
    x = E.tf32_round(torch.randn(M, Cc, device=dev))
    res = torch.randn(M, Cc, device=dev)
    p = E.Plan(dev)
    out = p.new_act(M, 1, 1, 1, Cc)
    p.mlp(E.Act(x, M, 1, 1, 1, Cc), E.pack_matrix(torch.randn(hid, Cc) / Cc ** 0.5, torch.randn(hid), dev),
          E.pack_matrix(torch.randn(Cc, hid) / hid ** 0.5, torch.randn(Cc), dev), out, residual=E.Act(res, M, 1, 1, 1, Cc))
    p.run()
    torch.cuda.synchronize()
    t = p.time_ops(iters=20)[0]
    print(f"{label:40s} M={M} C={Cc}: {t:.4f} ms  ({4.0 * M * Cc * hid / t / 1e9:.0f} TF/s)", flush=True)
    if hasattr(p.lib, "svx_mlp_prof_read"):   # SVX_MLP_PROFILE build: role timers of CTA 0, cycles per tile
        import ctypes
        buf = (ctypes.c_uint64 * 20)()
        p.lib.svx_mlp_prof_read(buf)
        nt = max(int(buf[7]), 1)
        names = ["mma:x_full", "mma:acc1_empty", "mma:w_full(fc1)", "mma:acc2_empty", "mma:h_full", "mma:w_full(fc2)",
                 "mma:total", "tiles", "epi:acc1_full", "epi:h_empty", "epi:r_full", "epi:acc2_full", "epi:total",
                 "epi:tmem_ld", "epi:bias+gelu", "epi:sts+fence", "epi:tail(incl. waits)"]
        print("   " + "  ".join(f"{nm}={int(buf[k]) / nt:.0f}" for k, nm in enumerate(names) if k != 7) + f"  (tiles={nt})")
