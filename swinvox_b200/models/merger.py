"""Drop-in for the reference's models/merger.py (:10-107): per-view 3-D conv stack producing voxel-wise scores,
softmax over the views and weighted fusion of the coarse volumes.  forward() replays graph.lower_merger."""
import torch
import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import PlanarInput, PlannedModule, mark_owned, owned_src, src_key


class Merger(PlannedModule):
    # MMA operand type of the 3x3x3 convolutions: "fp16" (default) is exact for the TF32-rounded activations this path
    # stores as long as |x| <= 65504; larger activations saturate and are reported by `saturated()`.  "tf32" keeps fp32's
    # exponent range at twice the tensor-core instructions.  Set before the first forward or call invalidate().
    slab_operands = "fp16"

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
            bound = isinstance(src, torch.Tensor) and tuple(src.shape) == (N * 34 ** 3, graph.RAW_CS)
            # the decoder's own output buffer (zero-bordered 34^3 x RAW_CS-channel rows), or a private one
            raw = E.Act(plan.hold(src) if bound else plan.zeros(N * 34 ** 3, graph.RAW_CS), N, 34, 34, 34, 16, 0, (1, 1, 1))
            coarse = PlanarInput(plan, coarse_volumes, (N, 32768))
            flag = plan.zeros(1, dtype=torch.int32)
            merged, weights = graph.lower_merger(plan, self, raw, coarse.buf, B, V, operands=self.slab_operands,
                                                 range_flag=flag)
            plan.range_flag = flag
            return plan, (raw, bound), coarse, merged, weights

        plan, (raw, bound), coarse, merged, weights = self._plan_for(
            (B, V, str(raw_features.device), src_key(raw_features), src_key(coarse_volumes), self.slab_operands), build)
        if not bound:   # foreign [B,V,9,32,32,32] tensor: re-layout into the interior (module-boundary path only)
            raw.view()[..., :9].copy_(E.tf32_round(raw_features.reshape(N, 9, 32, 32, 32)).permute(0, 2, 3, 4, 1))
        coarse.feed(coarse_volumes)
        plan.run(self.use_graph)
        self.last_volume_weights = weights.view(B, V, 32, 32, 32)   # pre-softmax scores (parity/debug)
        return mark_owned(merged.view(B, 32, 32, 32), merged)

    def saturated(self):
        """True if any forward since the plans were built fed the fp16-operand convolutions an activation beyond fp16's
        finite range (the result was computed from the saturated value).  Synchronises; meant for the end of an
        evaluation.  On True the module switches itself to "tf32" operands for the following forwards."""
        hit = any(int(e[0].range_flag.item()) != 0 for e in self._plans.values() if hasattr(e[0], "range_flag"))
        if hit:
            self.slab_operands = "tf32"
            self.invalidate()
        return hit
