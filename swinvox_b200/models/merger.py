"""Drop-in for the reference's models/merger.py (:10-107): per-view 3-D conv stack producing voxel-wise scores,
softmax over the views and weighted fusion of the coarse volumes.  forward() replays graph.lower_merger."""
import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import src_key, ChannelsLastInput, PlanarInput, PlannedModule, mark_owned


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
            raw = ChannelsLastInput(plan, raw_features, N, 9, 32768, 16, round_in=True)
            coarse = PlanarInput(plan, coarse_volumes, (N, 32768))
            merged, weights = graph.lower_merger(plan, self, E.Act(raw.buf, N, 32, 32, 32, 16), coarse.buf, B, V)
            return plan, raw, coarse, merged, weights

        plan, raw, coarse, merged, weights = self._plan_for((B, V, str(raw_features.device), src_key(raw_features), src_key(coarse_volumes)), build)
        raw.feed(raw_features)
        coarse.feed(coarse_volumes)
        plan.run(self.use_graph)
        self.last_volume_weights = weights.view(B, V, 32, 32, 32)   # pre-softmax scores (parity/debug)
        return mark_owned(merged.view(B, 32, 32, 32), merged)
