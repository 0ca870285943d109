"""Drop-in for the reference's models/cross_view_attention.py (:11-134): same constructor, parameters and
state_dict; forward() replays graph.lower_cva on libswinvox_b200."""
import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import src_key, ChannelsLastInput, PlannedModule, mark_owned


class CrossViewAttention(PlannedModule):
    def __init__(self, cfg, in_channels):
        super().__init__()
        net = cfg.NETWORK
        self.cfg, self.in_channels = cfg, in_channels
        self.num_heads = net.CROSS_ATT_NUM_HEADS
        self.reduced_channels = in_channels // net.CROSS_ATT_REDUCTION_RATIO
        self.attention_spatial_downsample_ratio = net.ATT_SPATIAL_DOWNSAMPLE_RATIO
        if self.reduced_channels % self.num_heads:
            raise AssertionError(f"reduced_channels ({self.reduced_channels}) must be divisible by num_heads ({self.num_heads})")
        self.head_dim = self.reduced_channels // self.num_heads
        r = self.attention_spatial_downsample_ratio
        self.downsample_qkv = nn.Conv2d(in_channels, in_channels, kernel_size=r, stride=r, groups=in_channels) if r > 1 else None
        self.qkv_conv = nn.Conv2d(in_channels, 3 * self.reduced_channels, kernel_size=1)
        self.softmax = nn.Softmax(dim=-1)
        self.proj_conv = nn.Conv2d(self.reduced_channels, in_channels, kernel_size=1)
        self.ffn = nn.Sequential(nn.Conv2d(in_channels, in_channels, kernel_size=1), nn.GELU(),
                                 nn.Conv2d(in_channels, in_channels, kernel_size=1))
        self.batch_norm = nn.BatchNorm2d(in_channels)
        self.dropout = nn.Dropout(0.1)

    def forward(self, x):
        self._guard(x)
        B, V, Cc, H, W = x.shape
        N = B * V

        def build():
            plan = E.Plan(x.device)
            inp = ChannelsLastInput(plan, x, N, Cc, H * W, Cc, round_in=True)
            out = graph.lower_cva(plan, self, E.Act(inp.buf, N, 1, H, W, Cc), B, V)
            return plan, inp, out

        plan, inp, out = self._plan_for((B, V, Cc, H, W, str(x.device), src_key(x)), build)
        inp.feed(x)
        plan.run(self.use_graph)
        return mark_owned(out.buf.view(B, V, H, W, Cc).permute(0, 1, 4, 2, 3), out.buf)
