"""Drop-in for the reference's models/swin_transformer.py (SwinTransformer wrapper, :11-94) with the timm
backbone 'swin_tiny_patch4_window7_224' (features_only) re-expressed as a parameter container in timm's key
layout.  forward() replays a libswinvox_b200 plan (graph.lower_swin); nothing here computes with PyTorch."""
import logging

import torch
import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import PlannedModule, mark_owned

_DEPTHS, _DIM, _WS = (2, 2, 6, 2), 96, 7


class _Holder(nn.Module):
    """parameter container; never called"""

    def forward(self, *a, **k):
        raise RuntimeError("swinvox_b200 parameter containers are not callable; use the owning module's forward()")


class _PatchEmbed(_Holder):
    def __init__(self):
        super().__init__()
        self.proj = nn.Conv2d(3, _DIM, kernel_size=4, stride=4)
        self.norm = nn.LayerNorm(_DIM)


class _Attn(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * _WS - 1) ** 2, dim // 32))
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _Mlp(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, 4 * dim), nn.Linear(4 * dim, dim)


class _Block(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.norm1, self.attn, self.norm2, self.mlp = nn.LayerNorm(dim), _Attn(dim), nn.LayerNorm(dim), _Mlp(dim)


class _Merge(_Holder):
    def __init__(self, dim_in):
        super().__init__()
        self.norm = nn.LayerNorm(4 * dim_in)
        self.reduction = nn.Linear(4 * dim_in, 2 * dim_in, bias=False)


class _Stage(_Holder):
    def __init__(self, s):
        super().__init__()
        dim = _DIM << s
        self.downsample = _Merge(dim // 2) if s else nn.Identity()
        self.blocks = nn.Sequential(*(_Block(dim) for _ in range(_DEPTHS[s])))


class _FeatureInfo:
    def __init__(self, chans):
        self._c = list(chans)

    def channels(self):
        return list(self._c)


class _SwinTFeatures(_Holder):
    """timm FeatureListNet layout: patch_embed, layers_0..layers_{last requested stage}"""

    def __init__(self, out_indices):
        super().__init__()
        idx = [i % 4 for i in out_indices]
        self.patch_embed = _PatchEmbed()
        for s in range(max(idx) + 1):
            self.add_module(f"layers_{s}", _Stage(s))
        self.feature_info = _FeatureInfo(_DIM << i for i in idx)
        for m in self.modules():   # timm's init: trunc_normal(.02) linears and bias tables
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, _Attn):
                nn.init.trunc_normal_(m.relative_position_bias_table, std=.02)


class SwinTransformer(PlannedModule):
    def __init__(self, cfg, in_channels=3, img_size=224, pretrained=True):
        super().__init__()
        self.cfg, self.img_size = cfg, img_size
        if img_size != 224:
            raise ValueError("swin_tiny_patch4_window7_224 is a fixed 224x224 model")
        if pretrained:
            logging.warning("swinvox_b200: pretrained Swin-T weights need network access; using random init "
                            "(load a checkpoint with load_state_dict).")
        stages = list(cfg.NETWORK.SWIN_T_STAGES)
        self.model = _SwinTFeatures(stages)
        old = self.model.patch_embed.proj
        self.model.patch_embed.proj = nn.Conv2d(in_channels, old.out_channels, old.kernel_size, old.stride, old.padding)
        if not pretrained:
            nn.init.xavier_uniform_(self.model.patch_embed.proj.weight)
            nn.init.zeros_(self.model.patch_embed.proj.bias)
        chans = self.model.feature_info.channels()
        self.out_channels = [chans[i] for i in range(len(stages))]
        self.out_spatial = [img_size // (4 * 2 ** i) for i in stages]
        self.layer_norm = nn.ModuleList(nn.LayerNorm([c, s, s]) for c, s in zip(self.out_channels, self.out_spatial))
        self.dropout = nn.Dropout(0.05)
        self.in_channels = in_channels

    def forward(self, x):
        """x: [N, in_channels, H, W] fp32.  Inputs that are not img_size x img_size are resized first (bilinear,
        align_corners=False), like models/swin_transformer.py:74-75."""
        self._guard(x)
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"SwinTransformer expects [N, {self.in_channels}, H, W], got {tuple(x.shape)}")
        if self.in_channels > 4:
            raise NotImplementedError("swinvox_b200 lowers patch embeddings of up to 4 input channels (16-byte pixels)")
        N, Cin, H, W = x.shape
        resize = (H, W) != (self.img_size, self.img_size)

        def build():
            plan = E.Plan(x.device)
            img = plan.empty(N, Cin, 224, 224)
            src = plan.empty(N, Cin, H, W) if resize else img
            if resize:
                plan.resize_bilinear(src, img, name="swin.input_resize")
            return plan, src, graph.lower_swin(plan, self, img, N)

        plan, src, feats = self._plan_for((N, H, W, str(x.device)), build)
        if x.data_ptr() != src.data_ptr():
            src.copy_(x)
        plan.run(self.use_graph)
        outs = [mark_owned(f.buf.view(N, f.H, f.W, f.C).permute(0, 3, 1, 2), f.buf) for f in feats]
        return outs if self.cfg.NETWORK.USE_SWIN_T_MULTI_STAGE else outs[-1]
