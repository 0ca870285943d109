"""Drop-in for the reference's models/refiner.py (:10-106): 3-D U-Net with a fully connected bottleneck,
[B,32,32,32] -> [B,32,32,32] logits.  forward() replays graph.lower_refiner."""
import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import src_key, PlanarInput, PlannedModule, mark_owned


class Refiner(PlannedModule):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        slope, bias = cfg.NETWORK.LEAKY_VALUE, cfg.NETWORK.TCONV_USE_BIAS

        def down(cin, cout):
            return nn.Sequential(nn.Conv3d(cin, cout, kernel_size=4, padding=2), nn.BatchNorm3d(cout), nn.LeakyReLU(slope),
                                 nn.MaxPool3d(kernel_size=2))

        def up(cin, cout):
            return nn.Sequential(nn.ConvTranspose3d(cin, cout, kernel_size=4, stride=2, bias=bias, padding=1),
                                 nn.BatchNorm3d(cout), nn.ReLU())

        self.layer1, self.layer2, self.layer3 = down(1, 32), down(32, 64), down(64, 128)
        self.layer4 = nn.Sequential(nn.Linear(8192, 2048), nn.ReLU())
        self.layer5 = nn.Sequential(nn.Linear(2048, 8192), nn.ReLU())
        self.layer6, self.layer7 = up(128, 64), up(64, 32)
        self.layer8 = nn.Sequential(nn.ConvTranspose3d(32, 1, kernel_size=4, stride=2, bias=bias, padding=1))

    def forward(self, coarse_volumes):
        self._guard(coarse_volumes)
        B = coarse_volumes.shape[0]
        if tuple(coarse_volumes.shape[1:]) != (32, 32, 32):
            raise ValueError(f"Refiner expects [B, 32, 32, 32] volumes, got {tuple(coarse_volumes.shape)}")

        def build():
            plan = E.Plan(coarse_volumes.device)
            vol = PlanarInput(plan, coarse_volumes, (B, 32768))
            return plan, vol, graph.lower_refiner(plan, self, vol.buf, B)

        plan, vol, out = self._plan_for((B, str(coarse_volumes.device), src_key(coarse_volumes)), build)
        vol.feed(coarse_volumes)
        plan.run(self.use_graph)
        return mark_owned(out.view(B, 32, 32, 32), out)
