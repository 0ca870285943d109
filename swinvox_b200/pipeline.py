"""The multi-view reconstruction forward chained exactly as the reference's evaluation loop does
(core/test.py:120-130,141-164): encoder -> decoder -> merger | mean -> refiner -> sigmoid/threshold/IoU, plus the
object-sharded data-parallel wrapper that replaces nn.DataParallel (core/test.py:72-76)."""
import torch
import torch.distributed as dist

from .metrics import VoxelMetrics
from .models import Decoder, Encoder, Merger, Refiner


class Reconstructor:
    def __init__(self, cfg, encoder=None, decoder=None, merger=None, refiner=None, device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.encoder = (encoder or Encoder(cfg)).eval().to(self.device)
        self.decoder = (decoder or Decoder(cfg)).eval().to(self.device)
        self.merger = (merger or Merger(cfg)).eval().to(self.device) if cfg.NETWORK.USE_MERGER else None
        self.refiner = (refiner or Refiner(cfg)).eval().to(self.device) if cfg.NETWORK.USE_REFINER else None
        self.metrics = VoxelMetrics(cfg.TEST.VOXEL_THRESH)

    def modules(self):
        return [m for m in (self.encoder, self.decoder, self.merger, self.refiner) if m is not None]

    def set_graph(self, on=True):
        for m in self.modules():
            m.use_graph = on

    def load_checkpoint(self, ckpt):
        """a reference checkpoint dict ({encoder,decoder,refiner,merger}_state_dict, keys possibly prefixed with
        DataParallel's `module.`; core/train.py:358-369)"""
        def strip(sd):
            return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        self.encoder.load_state_dict(strip(ckpt["encoder_state_dict"]))
        self.decoder.load_state_dict(strip(ckpt["decoder_state_dict"]))
        if self.merger is not None and "merger_state_dict" in ckpt:
            self.merger.load_state_dict(strip(ckpt["merger_state_dict"]))
        if self.refiner is not None and "refiner_state_dict" in ckpt:
            self.refiner.load_state_dict(strip(ckpt["refiner_state_dict"]))

    def input_buffer(self, B, V):
        return self.encoder.input_buffer(B, V, self.device)

    @torch.no_grad()
    def forward(self, images):
        """images [B,V,3,224,224] fp32 on the device -> refined occupancy logits [B,32,32,32]"""
        raw, gen = self.decoder(self.encoder(images))
        vol = self.merger(raw, gen) if self.merger is not None else gen.mean(dim=1)
        return self.refiner(vol) if self.refiner is not None else vol

    __call__ = forward

    @torch.no_grad()
    def evaluate(self, images, gt):
        """-> (logits [B,32,32,32], counts int32 [B,T,5]) ; use VoxelMetrics.scores(counts) for IoU / F-score"""
        logits = self.forward(images)
        return logits, self.metrics.counts(logits, gt)

    def num_launches(self):
        n = 1  # metrics
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

    @torch.no_grad()
    def evaluate_local(self, images_local, gt_local):
        """images_local: this rank's shard.  Returns gathered (logits [B,32,32,32], counts [B,T,5]) on every rank."""
        logits, counts = self.recon.evaluate(images_local, gt_local)
        if self.world == 1:
            return logits, counts
        all_logits = [torch.empty_like(logits) for _ in range(self.world)]
        all_counts = [torch.empty_like(counts) for _ in range(self.world)]
        dist.all_gather(all_logits, logits.contiguous(), group=self.group)
        dist.all_gather(all_counts, counts, group=self.group)
        return torch.cat(all_logits), torch.cat(all_counts)

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
        return logits[:B], counts[:B]
