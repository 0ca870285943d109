"""Drop-in for the reference's models/decoder.py (:11-99): [B,V,256,7,7] -> raw [B,V,9,32,32,32], coarse
[B,V,32,32,32].  forward() replays graph.lower_decoder (transposed convolutions as parity-class implicit GEMMs,
layer4+layer5+cat fused in one epilogue)."""
import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import src_key, ChannelsLastInput, PlannedModule, mark_owned


def _up(cin, cout, kernel, pad, bias):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, kernel_size=kernel, stride=2, bias=bias, padding=pad),
                         nn.BatchNorm3d(cout), nn.ReLU())


class Decoder(PlannedModule):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        bias = cfg.NETWORK.TCONV_USE_BIAS
        self.spatial_reduce = nn.AdaptiveAvgPool2d((2, 2))
        self.layer1 = _up(256, 128, (6, 4, 4), (2, 1, 1), bias)
        self.layer2 = _up(128, 64, 4, 1, bias)
        self.layer3 = _up(64, 32, 4, 1, bias)
        self.layer4 = _up(32, 8, 4, 1, bias)
        self.layer5 = nn.Sequential(nn.ConvTranspose3d(8, 1, kernel_size=1, bias=bias))

    def forward(self, image_features):
        self._guard(image_features)
        B, V, Cc, H, W = image_features.shape
        if (Cc, H, W) != (256, 7, 7):
            raise ValueError(f"Decoder expects [B, V, 256, 7, 7] features, got {tuple(image_features.shape)}")
        N = B * V

        def build():
            plan = E.Plan(image_features.device)
            inp = ChannelsLastInput(plan, image_features, N, 256, 49, 256, round_in=False)
            raw, coarse = graph.lower_decoder(plan, self, E.Act(inp.buf, N, 1, 7, 7, 256), N)
            return plan, inp, raw, coarse

        plan, inp, raw, coarse = self._plan_for((B, V, str(image_features.device), src_key(image_features)), build)
        inp.feed(image_features)
        plan.run(self.use_graph)
        raw_features = mark_owned(raw.buf.view(B, V, 34, 34, 34, graph.RAW_CS)[:, :, 1:33, 1:33, 1:33, :9].permute(0, 1, 5, 2, 3, 4), raw.buf)
        gen_volumes = mark_owned(coarse.view(B, V, 32, 32, 32), coarse)
        return raw_features, gen_volumes
