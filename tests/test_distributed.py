"""Multi-process path on CPU: world_size 2 over gloo.  The kernels are replaced by the test-only CPU twin
(tests/hostsim); what is checked is the host logic of the object-sharded data-parallel evaluation
(swinvox_b200/pipeline.py: shard selection, padding of an uneven batch, all_gather, trimming)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG_OVER = dict(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3], USE_CROSS_VIEW_ATTENTION=False)


def test_shard_objects_covers_batch_once():
    from swinvox_b200.pipeline import shard_objects
    for B in (1, 7, 8, 64, 257):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_objects(B, r, world)
                assert 0 <= lo <= hi <= B
                seen += list(range(lo, hi))
            assert seen == list(range(B))


def _build(cfg):
    from oracle import fixtures as FX
    from swinvox_b200.models import Decoder, Encoder, Merger, Refiner
    from swinvox_b200.pipeline import Reconstructor
    prod = FX.build(cfg, "calibrated", 0, dict(encoder=Encoder, decoder=Decoder, merger=Merger, refiner=Refiner))
    return Reconstructor(cfg, prod["encoder"], prod["decoder"], prod["merger"], prod["refiner"], device="cpu")


def _use_hostsim():
    from swinvox_b200 import _lib
    from swinvox_b200.models import _base
    import swinvox_b200.metrics as metrics
    _lib._lib = _lib.bind(os.path.join(ROOT, "tests", "hostsim", "libsvx_hostsim.so"))
    _base.require_device = lambda t: None


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    _use_hostsim()
    from oracle import fixtures as FX
    from oracle import modules as M
    from swinvox_b200.pipeline import DataParallelReconstructor
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = M.default_cfg(**CFG_OVER)
    rec = _build(cfg)
    images, gt = FX.structured_inputs(3, 1, seed=5), FX.seeded_gt(3)     # 3 objects over 2 ranks: uneven
    logits, counts = DataParallelReconstructor(rec).evaluate(images, gt)
    torch.save((logits, counts), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_matches_single_process(hostsim, tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert torch.equal(r0[0], r1[0]) and torch.equal(r0[1], r1[1]), "every rank must hold the same gathered result"
    # single-process evaluation of the same three objects
    from oracle import fixtures as FX
    from oracle import modules as M
    from swinvox_b200 import _lib
    from swinvox_b200.models import _base
    import swinvox_b200.metrics as metrics
    saved = (_lib._lib, _base.require_device)
    _use_hostsim()
    try:
        cfg = M.default_cfg(**CFG_OVER)
        rec = _build(cfg)
        images, gt = FX.structured_inputs(3, 1, seed=5), FX.seeded_gt(3)
        with torch.no_grad():
            logits, counts = rec.evaluate(images, gt)
        assert r0[0].shape == (3, 32, 32, 32) and r0[1].shape == (3, 4, 5)
        assert torch.allclose(r0[0], logits, rtol=0, atol=1e-6)
        assert torch.equal(r0[1], counts)
    finally:
        _lib._lib, _base.require_device = saved


def _eval_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    _use_hostsim()
    from oracle import modules as M
    from swinvox_b200.evaluate import BatchedEvaluator
    from test_evaluate import FakeRecon, make_data
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = M.default_cfg()
    tax, images, gt = make_data(7, 2, 11)
    mine = list(range(rank, 7, world))          # interleaved shards: the ranks meet the taxonomies in different orders
    ev = BatchedEvaluator(FakeRecon(cfg, "cpu"), 2, 2)
    for lo in range(0, len(mine), 2):
        idx = mine[lo:lo + 2]
        ev.submit([tax[i] for i in idx], images[idx], gt[idx])
    max_iou, rep = ev.finish(print_tables=False)
    torch.save((max_iou, rep), os.path.join(out_dir, f"eval{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_evaluator_reports_global_tables(hostsim, tmp_path):
    """BatchedEvaluator on two ranks with disjoint shards: after finish() both ranks hold the single-process result"""
    import numpy as np
    from oracle import eval_loop as OE
    from oracle import modules as M
    from test_evaluate import FakeRecon, make_data
    port = 31500 + os.getpid() % 2000
    mp.spawn(_eval_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "eval0.pt"), weights_only=False)
    r1 = torch.load(os.path.join(tmp_path, "eval1.pt"), weights_only=False)
    cfg = M.default_cfg()
    tax, images, gt = make_data(7, 2, 11)
    rec = FakeRecon(cfg, "cpu")
    refined = rec.refiner(rec.merger(None, images))
    ti, tf, mi, mf = OE.accumulate(tax, list(refined), list(gt), cfg.TEST.VOXEL_THRESH)
    for max_iou, rep in (r0, r1):
        assert rep["n_samples"] == 7 and abs(max_iou - float(np.max(mi))) < 1e-6
        np.testing.assert_allclose(rep["mean_iou"], mi, atol=1e-6)
        np.testing.assert_allclose(rep["mean_fscore"], mf, atol=1e-6)
        for tid in ti:
            assert rep["test_iou"][tid]["n_samples"] == ti[tid]["n_samples"]
            np.testing.assert_allclose(rep["test_iou"][tid]["iou"], ti[tid]["iou"], atol=1e-6)
