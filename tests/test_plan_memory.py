"""Activation-buffer reuse of engine.Plan (release / per-lane pools): host logic, run on the CPU twin; the GPU tier
re-checks the numbers of every pipeline test with reuse active (it is always on)."""
import torch

from oracle import fixtures as FX
from oracle import modules as M
from swinvox_b200 import engine as E
from swinvox_b200.models import Decoder, Encoder, Merger, Refiner
from util import dev  # noqa: F401


def _ptr(a):
    return a.buf.data_ptr()


def test_release_reuses_memory_in_the_same_lane_only(dev):
    p = E.Plan(dev)
    a = p.new_act(4, 1, 8, 8, 32)
    pa = _ptr(a)
    b = p.new_act(4, 1, 8, 8, 32)
    assert _ptr(b) != pa
    p.release(a)
    p.lane(1)
    c = p.new_act(4, 1, 8, 8, 32)          # another lane may overlap a's readers: fresh memory
    assert _ptr(c) not in (pa, _ptr(b))
    p.lane(0)
    d = p.new_act(4, 1, 8, 8, 16)          # same lane, smaller: best fit takes a's block
    assert _ptr(d) == pa and tuple(d.buf.shape) == (256, 16)
    p.release(a)                            # double release / release of a handed-out block: ignored
    e = p.new_act(4, 1, 8, 8, 32)
    assert _ptr(e) not in (pa, _ptr(b), _ptr(c))
    p.lane(1)
    p.release(c)
    p.join()                                # everything before the join precedes everything after it
    f = p.new_act(4, 1, 8, 8, 32)
    assert _ptr(f) == _ptr(c)


def test_zero_bordered_buffers_are_reused_by_identical_geometry_only(dev):
    p = E.Plan(dev)
    a = p.new_act(2, 1, 6, 6, 32, pad=(0, 1, 1))
    z = p.new_act(2, 1, 6, 6, 4, zero=True)     # relies on its zeros: never pooled
    pa, pz = _ptr(a), _ptr(z)
    p.release(a)
    p.release(z)
    b = p.new_act(2, 1, 8, 8, 32)                # plain buffers never take a bordered one (its border must stay zero)
    c = p.new_act(2, 1, 6, 6, 32, pad=(0, 1, 1))
    d = p.new_act(2, 1, 6, 6, 32, pad=(0, 1, 1))
    assert _ptr(b) not in (pa, pz) and _ptr(c) == pa and _ptr(d) not in (pa, pz)
    assert float(c.buf.abs().sum()) == 0.0


def test_bf16_and_fp32_share_blocks_by_size(dev):
    p = E.Plan(dev, dtype=torch.bfloat16)
    a = p.new_act(2, 1, 4, 4, 64)                # bf16: 4096 bytes
    pa = _ptr(a)
    p.release(a)
    b = p.new_act(2, 1, 4, 4, 32, dtype=torch.float32)
    assert _ptr(b) == pa and b.buf.dtype == torch.float32 and tuple(b.buf.shape) == (32, 32)


def test_encoder_plan_footprint(dev):
    """the Swin / ResNet branches live in a handful of buffers per resolution: fresh activation memory of one encoder
    plan stays below 40 MB per image (it was ~100 MB per image without reuse: B128 x V12 did not fit 180 GB)"""
    cfg = M.default_cfg()
    enc = FX.build(cfg, "calibrated", 0, dict(encoder=Encoder, decoder=Decoder, merger=Merger, refiner=Refiner))["encoder"].to(dev)
    images = FX.structured_inputs(1, 2, seed=3).to(dev)
    with torch.no_grad():
        enc(images)
    plan = next(iter(enc._plans.values()))[0]
    per_image = plan.bytes_allocated / 2 / 1e6
    assert plan.bytes_reused > plan.bytes_allocated, (plan.bytes_reused, plan.bytes_allocated)
    assert per_image < 40.0, per_image
