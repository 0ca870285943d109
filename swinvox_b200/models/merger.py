"""Drop-in for the reference's models/merger.py (:10-107): per-view 3-D conv stack producing voxel-wise scores,
softmax over the views and weighted fusion of the coarse volumes.  forward() replays graph.lower_merger."""
import torch
import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import PlanarInput, PlannedModule, mark_owned, owned_src, src_key


class Merger(PlannedModule):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        slope = cfg.NETWORK.LEAKY_VALUE

        def unit(cin, cout):
            return nn.Sequential(nn.Conv3d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm3d(cout), nn.LeakyReLU(slope))

        for i in range(1, 5):
            setattr(self, f"layer{i}", unit(9, 9))
        self.layer5 = unit(36, 9)
        self.layer6 = unit(9, 1)

    def forward(self, raw_features, coarse_volumes):
        self._guard(raw_features, coarse_volumes)
        B, V = raw_features.shape[:2]
        if tuple(raw_features.shape[2:]) != (9, 32, 32, 32) or tuple(coarse_volumes.shape) != (B, V, 32, 32, 32):
            raise ValueError("Merger expects raw_features [B,V,9,32,32,32] and coarse_volumes [B,V,32,32,32]")
        N = B * V

        def build():
            plan = E.Plan(raw_features.device)
            src = owned_src(raw_features)
            bound = isinstance(src, torch.Tensor) and tuple(src.shape) == (N * 34 ** 3, 32)
            # the decoder's own output buffer (zero-bordered 34^3 x 32-channel rows), or a private one
            raw = E.Act(plan.hold(src) if bound else plan.zeros(N * 34 ** 3, 32), N, 34, 34, 34, 16, 0, (1, 1, 1))
            coarse = PlanarInput(plan, coarse_volumes, (N, 32768))
            merged, weights = graph.lower_merger(plan, self, raw, coarse.buf, B, V)
            return plan, (raw, bound), coarse, merged, weights

        plan, (raw, bound), coarse, merged, weights = self._plan_for(
            (B, V, str(raw_features.device), src_key(raw_features), src_key(coarse_volumes)), build)
        if not bound:   # foreign [B,V,9,32,32,32] tensor: re-layout into the interior (module-boundary path only)
            raw.view()[..., :9].copy_(E.tf32_round(raw_features.reshape(N, 9, 32, 32, 32)).permute(0, 2, 3, 4, 1))
        coarse.feed(coarse_volumes)
        plan.run(self.use_graph)
        self.last_volume_weights = weights.view(B, V, 32, 32, 32)   # pre-softmax scores (parity/debug)
        return mark_owned(merged.view(B, 32, 32, 32), merged)
