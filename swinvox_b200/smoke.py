"""One small invocation of the hot path on cuda:0, checked against the oracle (used by __graft_entry__.smoke)."""
import torch


def run():
    from oracle import fixtures as FX
    from oracle import modules as M
    from swinvox_b200.models import Decoder, Encoder, Merger, Refiner
    from swinvox_b200.pipeline import Reconstructor
    cfg = M.default_cfg()
    B, V = 1, 2
    images, gt = FX.structured_inputs(B, V, seed=1234), FX.seeded_gt(B)
    ora = FX.build(cfg, "calibrated", 0)
    with torch.no_grad():
        ref = M.forward_pipeline(ora["encoder"], ora["decoder"], ora["merger"], ora["refiner"], images, cfg)
    prod = FX.build(cfg, "calibrated", 0, dict(encoder=Encoder, decoder=Decoder, merger=Merger, refiner=Refiner))
    rec = Reconstructor(cfg, prod["encoder"], prod["decoder"], prod["merger"], prod["refiner"], device="cuda:0")
    logits, counts = rec.evaluate(images.cuda(0), gt.cuda(0))
    torch.cuda.synchronize()
    got = logits.float().cpu()
    rel = (torch.linalg.vector_norm(got - ref) / torch.linalg.vector_norm(ref)).item()
    ref_counts, _, _ = M.voxel_metrics(ref, gt)
    dcount = (counts.cpu().long() - ref_counts).abs().max().item()
    print(f"smoke: B={B} V={V} rel_l2 vs oracle {rel:.2e}; max counter delta {dcount}; launches {rec.num_launches()}")
    assert rel < 1e-3, rel
    assert dcount <= 8, dcount
