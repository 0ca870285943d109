"""Batched evaluation driver against the restated accumulation / reporting of the reference's
test loop (oracle/eval_loop.py <- core/test.py:141-262): same per-taxonomy means, same overall means, same printed
tables, same losses -- with batches instead of single samples, a ragged last batch and interleaved taxonomies.
CPU tier: CPU twin of the metric kernel; gpu tier: the CUDA kernel + the overlapped copy-stream staging."""
import numpy as np
import pytest
import torch

from oracle import eval_loop as OE
from oracle import modules as M
from util import dev  # noqa: F401


class FakeRecon:
    """stands in for pipeline.Reconstructor: deterministic 'network' so the test isolates the driver"""

    def __init__(self, cfg, device):
        self.cfg, self.device = cfg, torch.device(device)
        self.encoder = lambda x: x
        self.decoder = lambda x: (None, x)
        self.merger = lambda raw, x: x.flatten(1)[:, :32768].reshape(-1, 32, 32, 32) * 6.0
        self.refiner = lambda v: v * 1.5 - 0.25

    def input_buffer(self, B, V):
        return torch.zeros(B, V, 3, 224, 224, device=self.device)

    def _gate(self, key):
        return True


def make_data(n, V, seed):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(n, V, 3, 224, 224, generator=g) * 2 - 1
    gt = (torch.rand(n, 32, 32, 32, generator=g) < 0.3).float()
    gt[2] = 0.0
    images[2] = -1.0            # empty prediction on empty ground truth: the union == 0 convention
    tax = ["02691156", "02828884", "02691156", "03001627", "02828884", "02691156", "03001627"][:n]
    return tax, images, gt


def test_batched_driver_matches_reference_loop(dev, monkeypatch):
    from swinvox_b200.evaluate import BatchedEvaluator
    cfg = M.default_cfg()
    th = cfg.TEST.VOXEL_THRESH
    n, V, B = 7, 2, 3
    tax, images, gt = make_data(n, V, 11)
    taxonomies = {"02691156": {"taxonomy_name": "aeroplane", "baseline": {"2-view": 0.5561, "1-view": 0.513}},
                  "02828884": {"taxonomy_name": "bench", "baseline": {"1-view": 0.421}},
                  "03001627": {"taxonomy_name": "chair"}}
    rec = FakeRecon(cfg, dev)
    seen = []
    ev = BatchedEvaluator(rec, B, V, taxonomies, on_batch=lambda lg, ct: seen.append((lg.clone(), ct.clone())))
    pin = (lambda t: t.pin_memory()) if dev == "cuda" else (lambda t: t)
    for lo in range(0, n, B):
        ev.submit(tax[lo:lo + B], pin(images[lo:lo + B]), pin(gt[lo:lo + B]))
    max_iou, rep = ev.finish(print_tables=False)

    # the reference's loop, one sample at a time, on the same logits
    merged = rec.merger(None, images)
    refined = rec.refiner(merged)
    ti, tf, mi, mf = OE.accumulate(tax, list(refined), list(gt), th)
    assert rep["n_samples"] == n
    # every batch's voxels reached the host consumer, in order, without the padding of the ragged last batch
    assert [t[0].shape[0] for t in seen] == [3, 3, 1]
    assert torch.equal(torch.cat([t[0] for t in seen]).cpu(), refined.cpu())
    assert list(rep["test_iou"]) == list(ti)                      # first-seen taxonomy order, as the reference's dicts
    for tid in ti:
        assert rep["test_iou"][tid]["n_samples"] == ti[tid]["n_samples"]
        np.testing.assert_allclose(rep["test_iou"][tid]["iou"], ti[tid]["iou"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(rep["test_fscore"][tid]["fscore"], tf[tid]["fscore"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(rep["mean_iou"], mi, atol=1e-6)
    np.testing.assert_allclose(rep["mean_fscore"], mf, atol=1e-6)
    assert abs(max_iou - float(np.max(mi))) < 1e-6
    losses = [OE.sample_losses(merged[i:i + 1], refined[i:i + 1], gt[i:i + 1]) for i in range(n)]
    assert abs(rep["encoder_loss"] - np.mean([a for a, _ in losses])) < 2e-4
    assert abs(rep["refiner_loss"] - np.mean([b for _, b in losses])) < 2e-4
    assert ev.tables(rep) == OE.tables(ti, tf, mi, mf, taxonomies, th, V)


def test_tables_layout_is_the_references():
    """the literal layout of core/test.py:222-262 for a two-taxonomy result"""
    th = [0.2, 0.3, 0.4, 0.5]
    ti = {"a": {"n_samples": 2, "iou": np.array([0.5, 0.25, 0.125, 0.0625])}}
    tf = {"a": {"n_samples": 2, "fscore": np.array([0.1, 0.2, 0.3, 0.4])}}
    txt = OE.tables(ti, tf, np.array([0.5, 0.25, 0.125, 0.0625]), np.array([0.1, 0.2, 0.3, 0.4]),
                    {"a": {"taxonomy_name": "sofa", "baseline": {"3-view": 0.7}}}, th, 3)
    lines = txt.split("\n")
    assert lines[0] == '============================ TEST RESULTS (IoU) ============================'
    assert lines[1] == 'Taxonomy\t#Sample\tBaseline\tt=0.20\tt=0.30\tt=0.40\tt=0.50\t'
    assert lines[2] == 'sofa    \t2\t0.7000\t\t0.5000\t0.2500\t0.1250\t0.0625\t'
    assert lines[3] == 'Overall \t\t\t\t0.5000\t0.2500\t0.1250\t0.0625\t'
    assert lines[4] == '' and lines[5] == '========================== TEST RESULTS (F-score) =========================='
    assert lines[7] == 'sofa    \t2\tN/a\t\t0.1000\t0.2000\t0.3000\t0.4000\t'
