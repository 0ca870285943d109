"""The multi-view reconstruction forward chained exactly as the reference's evaluation loop does
(core/test.py:120-130,141-164): encoder -> decoder -> merger | mean -> refiner -> sigmoid/threshold/IoU, plus the
object-sharded data-parallel wrapper that replaces nn.DataParallel (core/test.py:72-76)."""
import re

import torch
import torch.distributed as dist

from . import engine as E
from .metrics import VoxelMetrics
from .models import Decoder, Encoder, Merger, Refiner
from .models import _base
from .models._base import PlanarInput, mark_owned, src_key


class ViewMean:
    """torch.mean(generated_volume, dim=1) of core/test.py:125-126 (merger off, or before EPOCH_START_USE_MERGER) as one
    launch of the merger's fusion kernel with uniform weights; binds to the decoder's coarse buffer when chained."""

    def __init__(self):
        self._plans = {}

    def __call__(self, coarse_volumes):
        _base.require_device(coarse_volumes)
        if coarse_volumes.dtype != torch.float32:
            raise TypeError(f"expected float32 volumes, got {coarse_volumes.dtype}")
        B, V = coarse_volumes.shape[:2]
        P = coarse_volumes[0, 0].numel()
        key = (B, V, P, str(coarse_volumes.device), src_key(coarse_volumes))
        if key not in self._plans:
            if len(self._plans) > 8:
                self._plans.clear()
            plan = E.Plan(coarse_volumes.device)
            coarse = PlanarInput(plan, coarse_volumes, (B * V, P))
            out = plan.empty(B, P)
            plan.merger_fuse(None, coarse.buf, out, B, V, P, name="view_mean")
            self._plans[key] = (plan, coarse, out)
        plan, coarse, out = self._plans[key]
        coarse.feed(coarse_volumes)
        plan.run()
        return mark_owned(out.view(B, *coarse_volumes.shape[2:]), out)


_OLD_SINGLE_LN = re.compile(r"^swin_transformer\.layer_norm\.(weight|bias)$")


def adapt_state_dict(sd):
    """Key adapter for reference checkpoints: strips DataParallel's `module.` prefix (core/train.py:358-369 saves the
    wrapped modules) and maps the un-indexed wrapper LayerNorm of single-stage checkpoints written before
    models/swin_transformer.py:64-67 became a ModuleList (notebook cell 68: `swin_transformer.layer_norm.weight`) onto
    `swin_transformer.layer_norm.0.*`.  Anything else goes to load_state_dict(strict=True) untouched, which names the
    missing / unexpected keys of checkpoints from other architecture revisions."""
    out = {}
    for k, v in sd.items():
        k = k[7:] if k.startswith("module.") else k
        m = _OLD_SINGLE_LN.match(k)
        if m:
            k = f"swin_transformer.layer_norm.0.{m.group(1)}"
        out[k] = v
    return out


class Reconstructor:
    """Outputs: `forward` / `evaluate` return FRESH tensors (like the reference's modules) unless `zero_copy=True`, in
    which case they are views of plan-owned buffers that the next call with the same (B, V) overwrites."""

    def __init__(self, cfg, encoder=None, decoder=None, merger=None, refiner=None, device="cuda", zero_copy=False,
                 dtype="tf32"):
        """dtype: "tf32" (default) or "bf16" -- the encoder's storage / operand type (Encoder.compute_dtype); decoder,
        merger and refiner compute in fp32 / TF32 either way"""
        self.cfg = cfg
        self.device = torch.device(device)
        self.zero_copy = zero_copy
        self.encoder = (encoder or Encoder(cfg)).eval().to(self.device)
        self.encoder.compute_dtype = dtype
        self.decoder = (decoder or Decoder(cfg)).eval().to(self.device)
        self.merger = (merger or Merger(cfg)).eval().to(self.device) if cfg.NETWORK.USE_MERGER else None
        self.refiner = (refiner or Refiner(cfg)).eval().to(self.device) if cfg.NETWORK.USE_REFINER else None
        self.view_mean = ViewMean()
        self.metrics = VoxelMetrics(cfg.TEST.VOXEL_THRESH)
        self.epoch_idx = None   # set by load_checkpoint; None = no epoch gating (weights loaded by hand)

    def _gate(self, key):
        """core/test.py:123,129: a stage runs only once the checkpoint's epoch reached cfg.TRAIN.EPOCH_START_USE_<stage>"""
        train = getattr(self.cfg, "TRAIN", None)
        start = getattr(train, key, None) if train is not None else None
        return self.epoch_idx is None or start is None or self.epoch_idx >= start

    def modules(self):
        return [m for m in (self.encoder, self.decoder, self.merger, self.refiner) if m is not None]

    def set_graph(self, on=True):
        for m in self.modules():
            m.use_graph = on

    def load_checkpoint(self, ckpt):
        """a reference checkpoint dict ({encoder,decoder,refiner,merger}_state_dict, keys possibly prefixed with
        DataParallel's `module.`; core/train.py:358-369)"""
        # like core/test.py:82-89: a missing merger / refiner state dict raises KeyError when the cfg uses that stage
        self.encoder.load_state_dict(adapt_state_dict(ckpt["encoder_state_dict"]))
        self.decoder.load_state_dict(adapt_state_dict(ckpt["decoder_state_dict"]))
        if self.refiner is not None:
            self.refiner.load_state_dict(adapt_state_dict(ckpt["refiner_state_dict"]))
        if self.merger is not None:
            self.merger.load_state_dict(adapt_state_dict(ckpt["merger_state_dict"]))
        self.epoch_idx = ckpt.get("epoch_idx")

    def input_buffer(self, B, V):
        return self.encoder.input_buffer(B, V, self.device)

    @torch.no_grad()
    def forward(self, images):
        """images [B,V,3,224,224] fp32 on the device -> refined occupancy logits [B,32,32,32]"""
        vol = self._forward_views(images)
        return vol if self.zero_copy else vol.clone()

    def _forward_views(self, images):
        raw, gen = self.decoder(self.encoder(images))
        if self.merger is not None and self._gate("EPOCH_START_USE_MERGER"):
            vol = self.merger(raw, gen)
        else:
            vol = self.view_mean(gen)
        if self.refiner is not None and self._gate("EPOCH_START_USE_REFINER"):
            vol = self.refiner(vol)
        return vol

    __call__ = forward

    @torch.no_grad()
    def evaluate(self, images, gt):
        """-> (logits [B,32,32,32], counts int32 [B,T,5]) ; use VoxelMetrics.scores(counts) for IoU / F-score"""
        logits = self._forward_views(images)
        counts = self.metrics.counts(logits, gt)
        return (logits, counts) if self.zero_copy else (logits.clone(), counts.clone())

    def num_launches(self):
        n = 1  # metrics
        n += sum(e[0].num_launches for e in self.view_mean._plans.values())
        for m in self.modules():
            for entry in m._plans.values():
                n += entry[0].num_launches
        return n


def shard_objects(B, rank, world):
    """objects [lo, hi) owned by `rank`: the views of one object stay together (CVA / merger couple them)"""
    per = (B + world - 1) // world
    lo = min(rank * per, B)
    return lo, min(lo + per, B)


class DataParallelReconstructor:
    """One process per GPU, weights resident per rank, objects sharded; the only exchange is the final NCCL
    all_gather of logits and IoU counters (replaces nn.DataParallel's per-forward broadcast + gather)."""

    def __init__(self, recon, group=None):
        self.recon = recon
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._slot = 0
        self._work = [None, None]
        self._bufs = {}

    def _slot_buffers(self, s, B, P, C, device):
        key = (s, B, P, C, str(device))
        if key not in self._bufs:
            packed = torch.empty(B, P + C, dtype=torch.float32, device=device)
            out = torch.empty(self.world * B, P + C, dtype=torch.float32, device=device)
            self._bufs[key] = (packed, out)
        return self._bufs[key]

    @torch.no_grad()
    def evaluate_local(self, images_local, gt_local, wait=True):
        """images_local: this rank's shard.  Returns the gathered (logits [R*B,32,32,32], counts int32 [R*B,T,5]) on every
        rank (views of a preallocated gather buffer, valid until the call after next).
        The exchange is ONE all_gather_into_tensor of a packed [B, 32^3 + 5T] buffer (logits | counter bits).  With
        wait=False the gather is only enqueued (it runs on NCCL's stream while this rank's next forward computes); call
        `flush()` -- or the next-but-one evaluate_local -- before reading the returned tensors."""
        logits, counts = self.recon.evaluate(images_local, gt_local)
        if self.world == 1:
            return logits, counts
        B, P = logits.shape[0], logits[0].numel()
        Cn = counts[0].numel()
        s = self._slot
        self._slot ^= 1
        if self._work[s] is not None:      # the slot's previous gather (two calls ago) must have drained
            self._work[s].wait()
        packed, out = self._slot_buffers(s, B, P, Cn, logits.device)
        packed[:, :P].copy_(logits.reshape(B, P))
        packed[:, P:].copy_(counts.reshape(B, Cn).view(torch.float32))
        self._work[s] = dist.all_gather_into_tensor(out, packed, group=self.group, async_op=True)
        if wait:
            self._work[s].wait()
            self._work[s] = None
        R = self.world
        return (out[:, :P].unflatten(1, tuple(logits.shape[1:])),
                out[:, P:].view(torch.int32).unflatten(1, tuple(counts.shape[1:])))

    def flush(self):
        """the current stream waits for every gather still in flight"""
        for s in (0, 1):
            if self._work[s] is not None:
                self._work[s].wait()
                self._work[s] = None

    @torch.no_grad()
    def evaluate(self, images, gt):
        """images [B,V,3,224,224], gt [B,32,32,32]: the GLOBAL batch, identical on every rank.  Each rank evaluates its
        shard of objects (the last shards are padded by repeating the final object so every rank runs the same shape)
        and the gathered results are trimmed back to B objects."""
        B = images.shape[0]
        per = (B + self.world - 1) // self.world
        idx = torch.arange(self.rank * per, (self.rank + 1) * per).clamp_(max=B - 1)
        dev = self.recon.device
        logits, counts = self.evaluate_local(images[idx].to(dev), gt[idx].to(dev))
        return logits[:B].contiguous(), counts[:B].contiguous()   # fresh tensors (the gather buffer is reused)
