import torch, sys, os
sys.path.insert(0, os.getcwd())
from swinvox_b200 import engine as E
dev = "cuda"
for M in (128, 1000):
    C = 192; hid = 768
    torch.manual_seed(1)
    x = E.tf32_round(torch.randn(M, C)); res = torch.randn(M, C); b2 = torch.randn(C)
    w1, b1 = torch.randn(hid, C) / C ** 0.5, torch.randn(hid) * 0.5
    w2 = torch.randn(C, hid) / hid ** 0.5
    p = E.Plan(dev)
    out = p.new_act(M, 1, 1, 1, C)
    p.mlp(E.Act(x.to(dev), M, 1, 1, 1, C), E.pack_matrix(w1, b1, dev), E.pack_matrix(w2, b2, dev), out, residual=E.Act(res.to(dev), M, 1, 1, 1, C))
    p.run(); torch.cuda.synchronize()
    h = torch.nn.functional.gelu(x.double() @ E.tf32_round(w1).double().t() + b1.double())
    ref = E.tf32_round(h.float()).double() @ E.tf32_round(w2).double().t() + b2.double() + res.double()
    got = out.view().reshape(M, C).cpu().double()
    err = (got - ref).abs()
    print("M", M, "nan", torch.isnan(got).sum().item(), "max err", err.max().item(), "bad rows", (err.max(1).values > 1e-2).nonzero().flatten()[:10].tolist(),
          "bad cols", (err.max(0).values > 1e-2).nonzero().flatten()[:24].tolist())
