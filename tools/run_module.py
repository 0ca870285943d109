"""Runs one module of the pipeline a few times on random data (for ncu / per-op timing on the GPU box).

    python tools/run_module.py merger|decoder|refiner|encoder|all [B] [V] [iters]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from swinvox_b200 import config as svx_config  # noqa: E402
from swinvox_b200.pipeline import Reconstructor  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
V = int(sys.argv[3]) if len(sys.argv) > 3 else 3
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
cfg = svx_config.make_cfg()
rec = Reconstructor(cfg, device="cuda")
images = torch.rand(B, V, 3, 224, 224, device="cuda") * 2 - 1
with torch.no_grad():
    if os.environ.get("SVX_ISOLATE"):   # random stand-ins for the upstream modules (ncu captures only `which`)
        feat = torch.randn(B, V, 256, 7, 7, device="cuda")
        raw, gen = torch.randn(B, V, 9, 32, 32, 32, device="cuda"), torch.rand(B, V, 32, 32, 32, device="cuda")
        vol = torch.rand(B, 32, 32, 32, device="cuda")
    else:
        feat = rec.encoder(images)
        raw, gen = rec.decoder(feat)
        vol = rec.merger(raw, gen)
        out = rec.refiner(vol)
    torch.cuda.synchronize()
    fns = {"encoder": lambda: rec.encoder(images), "decoder": lambda: rec.decoder(feat),
           "merger": lambda: rec.merger(raw, gen), "refiner": lambda: rec.refiner(vol),
           "all": lambda: rec.forward(images)}
    fn = fns[which]
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{which} B={B} V={V}: {a.elapsed_time(b) / iters:.3f} ms per call")
    mod = getattr(rec, which, None)
    if mod is not None:
        for entry in mod._plans.values():
            plan = entry[0]
            for nm, t in zip(plan.op_names, plan.time_ops(iters=3)):
                print(f"  {nm:40s} {t:8.4f} ms")
