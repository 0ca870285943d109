"""Drop-in for the reference's models/encoder.py (:15-164): ResNet-50 trunk + Swin-T + optional cross-view
attention + fusion convs, [B,V,3,224,224] -> [B,V,256,7,7].  Parameters keep the reference state_dict layout
(resnet.*, swin_transformer.*, resnet_reduce.*, swin_stage_reduces.*, swin_downsamples.*, cross_view_attention.*,
fusion_layer.*, layer1-3.*); forward() replays graph.lower_encoder on libswinvox_b200."""
import logging

import torch.nn as nn

from .. import engine as E
from .. import graph
from ._base import PlannedModule, mark_owned
from .cross_view_attention import CrossViewAttention
from .swin_transformer import SwinTransformer


def _conv_bn_relu(cin, cout, stride):
    return [nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]


class Encoder(PlannedModule):
    # "tf32" (default: fp32 activations, TF32 tensor cores, the reference's fp32 numerics within rtol 1e-3) or "bf16"
    # (bf16 activation storage and tensor-core operands, fp32 accumulation / LayerNorm / softmax statistics; BASELINE
    # configs[2]).  The module's input and output stay fp32 either way.  Set before the first forward.
    compute_dtype = "tf32"

    def __init__(self, cfg):
        super().__init__()
        import torchvision
        self.cfg = cfg
        net = cfg.NETWORK
        logging.info("swinvox_b200: ResNet-50 / Swin-T start from random init (pretrained weights need network).")
        trunk = torchvision.models.resnet50(weights=None)
        self.resnet = nn.Sequential(*list(trunk.children())[:7])          # conv1 .. layer3 (parameter container)
        self.swin_transformer = SwinTransformer(cfg, in_channels=3, img_size=224, pretrained=True)
        self.resnet_reduce = nn.Conv2d(1024, 256, kernel_size=1)
        if net.USE_SWIN_T_MULTI_STAGE:
            self.swin_stage_reduces = nn.ModuleList(nn.Conv2d(c, 256, kernel_size=1) for c in self.swin_transformer.out_channels)
            depth = {0: 3, 1: 2, 2: 1}   # stride-2 convs needed to bring stage i down to 7x7
            self.swin_downsamples = nn.ModuleList(
                nn.Sequential(*[m for _ in range(depth[i]) for m in _conv_bn_relu(256, 256, 2)]) if i in depth else nn.Identity()
                for i in net.SWIN_T_STAGES)
        else:
            self.swin_reduce = nn.Conv2d(768, 256, kernel_size=1)
        self.cross_view_attention = CrossViewAttention(cfg, in_channels=512) if net.USE_CROSS_VIEW_ATTENTION else None
        self.fusion_layer = nn.Sequential(*_conv_bn_relu(512, 256, 1))
        self.layer1 = nn.Sequential(*_conv_bn_relu(256, 256, 1))
        self.layer2 = nn.Sequential(*_conv_bn_relu(256, 256, 1))
        self.layer3 = nn.Sequential(*_conv_bn_relu(256, 256, 1))

    def input_buffer(self, B, V, device):
        """the plan's own [B,V,3,224,224] staging tensor: write images into it (e.g. an H2D copy) and pass it to
        forward() to skip the device-to-device input copy"""
        import torch
        with torch.no_grad():
            plan, img, _ = self._get_plan(B, V, torch.device(device))
        return img.view(B, V, 3, 224, 224)

    def _get_plan(self, B, V, device):
        if self.compute_dtype not in ("tf32", "bf16"):
            raise ValueError(f"Encoder.compute_dtype must be 'tf32' or 'bf16', got {self.compute_dtype!r}")

        def build():
            import torch
            plan = E.Plan(device, dtype=torch.bfloat16 if self.compute_dtype == "bf16" else torch.float32)
            img = plan.empty(B * V, 3, 224, 224)
            return plan, img, graph.lower_encoder(plan, self, img, B, V)

        return self._plan_for((B, V, str(device), self.compute_dtype), build)

    def attention_saturated(self):
        """True if a window-attention V value exceeded fp16's finite range (65504) in any forward since the plans were
        built: with fp32 / TF32 storage the P V product of the attention kernel runs on fp16 operands, which is exact for
        this path's TF32-rounded values only inside that range.  Synchronises; meant for the end of an evaluation."""
        return any(int(e[0].taps["attn_range_flag"].item()) != 0 for e in self._plans.values()
                   if "attn_range_flag" in e[0].taps)

    def forward(self, rendering_images):
        self._guard(rendering_images)
        B, V, Cc, H, W = rendering_images.shape
        if (Cc, H, W) != (3, 224, 224):
            raise ValueError(f"Encoder expects [B, V, 3, 224, 224] images, got {tuple(rendering_images.shape)}")
        plan, img, out = self._get_plan(B, V, rendering_images.device)
        if rendering_images.data_ptr() != img.data_ptr():
            img.copy_(rendering_images.reshape(B * V, 3, 224, 224))
        plan.run(self.use_graph)
        return mark_owned(out.buf.view(B, V, 7, 7, 256).permute(0, 1, 4, 2, 3), out.buf)
