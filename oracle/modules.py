"""ORACLE (test infrastructure, not product code): CPU/PyTorch-fp32 restatement of the SwinVox forward path.

Each class restates one reference module with the SAME state_dict layout, so weights move freely between
the reference (/root/reference/models/*.py), this oracle and the product (swinvox_b200/models/*.py):

    RefEncoder              <- models/encoder.py:15-164
    RefSwinTransformer      <- models/swin_transformer.py:11-94      (timm backbone: oracle/swin_t.py)
    RefCrossViewAttention   <- models/cross_view_attention.py:11-134
    RefDecoder              <- models/decoder.py:11-99
    RefMerger               <- models/merger.py:10-107
    RefRefiner              <- models/refiner.py:10-106
    voxel_metrics           <- core/test.py:141-164

Pinned against the real reference modules (imported from /root/reference in the build container) by
oracle/make_golden.py, which also writes the fixtures under tests/golden/.  Only tests/, smoke() and
bench.py's CPU-baseline leg may import this package.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import swin_t


class AttrDict(dict):
    """stand-in for easydict (not installed): cfg.NETWORK.X attribute access over nested dicts"""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __setattr__(self, k, v):
        self[k] = v


def default_cfg(**network_overrides):
    """the keys of config.py:83-94,132 the hot path reads, with the reference defaults"""
    net = dict(LEAKY_VALUE=0.2, TCONV_USE_BIAS=False, USE_REFINER=True, USE_MERGER=True,
               USE_SWIN_T_MULTI_STAGE=True, SWIN_T_STAGES=[0, 1, 2, 3], USE_CROSS_VIEW_ATTENTION=True,
               CROSS_ATT_REDUCTION_RATIO=4, ATT_SPATIAL_DOWNSAMPLE_RATIO=2, CROSS_ATT_NUM_HEADS=4)
    net.update(network_overrides)
    return AttrDict(NETWORK=AttrDict(net), TEST=AttrDict(VOXEL_THRESH=[0.2, 0.3, 0.4, 0.5]),
                    CONST=AttrDict(IMG_W=224, IMG_H=224))


def _cbr2d(cin, cout, k, stride=1, pad=0):
    return [nn.Conv2d(cin, cout, k, stride, pad), nn.BatchNorm2d(cout), nn.ReLU()]


class RefSwinTransformer(nn.Module):
    def __init__(self, cfg, in_channels=3, img_size=224, pretrained=True):
        super().__init__()
        self.cfg, self.img_size = cfg, img_size
        stages = cfg.NETWORK.SWIN_T_STAGES
        self.model = swin_t.create_model("swin_tiny_patch4_window7_224", pretrained=False, features_only=True,
                                         out_indices=stages)
        old = self.model.patch_embed.proj
        self.model.patch_embed.proj = nn.Conv2d(in_channels, old.out_channels, old.kernel_size, old.stride, old.padding)
        if not pretrained:  # swin_transformer.py:50-54
            nn.init.xavier_uniform_(self.model.patch_embed.proj.weight)
            nn.init.zeros_(self.model.patch_embed.proj.bias)
        chans = self.model.feature_info.channels()
        self.out_channels = [chans[i] for i in range(len(stages))]
        self.out_spatial = [img_size // (4 * 2 ** i) for i in stages]
        self.layer_norm = nn.ModuleList(nn.LayerNorm([c, s, s]) for c, s in zip(self.out_channels, self.out_spatial))
        self.dropout = nn.Dropout(0.05)

    def forward(self, x):
        if tuple(x.shape[-2:]) != (self.img_size, self.img_size):
            x = F.interpolate(x, size=(self.img_size, self.img_size), mode="bilinear", align_corners=False)
        outs = [self.dropout(ln(f.permute(0, 3, 1, 2))) for f, ln in zip(self.model(x), self.layer_norm)]
        return outs if self.cfg.NETWORK.USE_SWIN_T_MULTI_STAGE else outs[-1]


class RefCrossViewAttention(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        net = cfg.NETWORK
        self.cfg, self.in_channels = cfg, in_channels
        self.num_heads = net.CROSS_ATT_NUM_HEADS
        self.reduced_channels = in_channels // net.CROSS_ATT_REDUCTION_RATIO
        self.ratio = net.ATT_SPATIAL_DOWNSAMPLE_RATIO
        assert self.reduced_channels % self.num_heads == 0
        self.head_dim = self.reduced_channels // self.num_heads
        self.downsample_qkv = (nn.Conv2d(in_channels, in_channels, self.ratio, self.ratio, groups=in_channels)
                               if self.ratio > 1 else None)
        self.qkv_conv = nn.Conv2d(in_channels, 3 * self.reduced_channels, 1)
        self.proj_conv = nn.Conv2d(self.reduced_channels, in_channels, 1)
        self.ffn = nn.Sequential(nn.Conv2d(in_channels, in_channels, 1), nn.GELU(), nn.Conv2d(in_channels, in_channels, 1))
        self.batch_norm = nn.BatchNorm2d(in_channels)
        self.dropout = nn.Dropout(0.1)

    def forward(self, x):
        B, V, Cc, H, W = x.shape
        flat = x.reshape(B * V, Cc, H, W)
        small = self.downsample_qkv(flat) if self.downsample_qkv is not None else flat
        h, w = small.shape[-2:]
        R, nh, hd = self.reduced_channels, self.num_heads, self.head_dim
        q, k, v = self.qkv_conv(small).split(R, dim=1)
        q = q.reshape(B, V, nh, hd * h * w).transpose(1, 2)               # [B, nh, V, L]
        k = k.reshape(B, V, nh, hd * h * w).transpose(1, 2)
        v = v.reshape(B, V, nh, hd * h * w).transpose(1, 2)
        attn = torch.softmax(q @ k.transpose(-1, -2) / (hd * V) ** 0.5, dim=-1)   # over the view axis
        y = (attn @ v).transpose(1, 2).reshape(B * V, R, h, w)
        y = self.proj_conv(y)
        if self.downsample_qkv is not None:
            y = F.interpolate(y, size=(H, W), mode="bilinear", align_corners=False)
        y = y + flat                          # residual, then FFN *replaces* (no second residual)
        y = self.dropout(self.batch_norm(self.ffn(y)))
        return y.reshape(B, V, Cc, H, W)


class RefEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        import torchvision
        self.cfg = cfg
        net = cfg.NETWORK
        trunk = torchvision.models.resnet50(weights=None)   # the reference asks for pretrained weights (network)
        self.resnet = nn.Sequential(*list(trunk.children())[:7])
        self.swin_transformer = RefSwinTransformer(cfg, in_channels=3, img_size=224, pretrained=True)
        self.resnet_reduce = nn.Conv2d(1024, 256, 1)
        if net.USE_SWIN_T_MULTI_STAGE:
            self.swin_stage_reduces = nn.ModuleList(nn.Conv2d(c, 256, 1) for c in self.swin_transformer.out_channels)
            chain = lambda n: nn.Sequential(*[m for _ in range(n) for m in _cbr2d(256, 256, 3, 2, 1)]) if n else nn.Identity()
            self.swin_downsamples = nn.ModuleList(chain({0: 3, 1: 2, 2: 1}.get(i, 0)) for i in net.SWIN_T_STAGES)
        else:
            self.swin_reduce = nn.Conv2d(768, 256, 1)
        self.cross_view_attention = RefCrossViewAttention(cfg, 512) if net.USE_CROSS_VIEW_ATTENTION else None
        self.fusion_layer = nn.Sequential(*_cbr2d(512, 256, 3, 1, 1))
        self.layer1 = nn.Sequential(*_cbr2d(256, 256, 3, 1, 1))
        self.layer2 = nn.Sequential(*_cbr2d(256, 256, 3, 1, 1))
        self.layer3 = nn.Sequential(*_cbr2d(256, 256, 3, 1, 1))

    def forward(self, images, taps=None):
        """taps: optional dict that receives the intermediate tensors (stage-boundary goldens)"""
        B, V = images.shape[:2]
        img = images.reshape(B * V, *images.shape[2:])
        res = F.avg_pool2d(self.resnet_reduce(self.resnet(img)), 2, 2)
        sw = self.swin_transformer(img)
        if self.cfg.NETWORK.USE_SWIN_T_MULTI_STAGE:
            parts = [down(red(f)) for f, red, down in zip(sw, self.swin_stage_reduces, self.swin_downsamples)]
            sw_sum = torch.stack(parts).sum(0)
        else:
            sw_sum = self.swin_reduce(sw)
        cat = torch.cat((res, sw_sum), 1).reshape(B, V, 512, 7, 7)
        att = self.cross_view_attention(cat) if self.cfg.NETWORK.USE_CROSS_VIEW_ATTENTION else cat
        y = self.layer3(self.layer2(self.layer1(self.fusion_layer(att.reshape(B * V, 512, 7, 7)))))
        if taps is not None:
            taps.update(resnet=res, swin=sw, swin_sum=sw_sum, pre_cva=cat, post_cva=att)
        return y.reshape(B, V, 256, 7, 7)


def _tbr3d(cin, cout, k, pad, bias):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, k, 2, pad, bias=bias), nn.BatchNorm3d(cout), nn.ReLU())


class RefDecoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        b = cfg.NETWORK.TCONV_USE_BIAS
        self.spatial_reduce = nn.AdaptiveAvgPool2d((2, 2))
        self.layer1 = _tbr3d(256, 128, (6, 4, 4), (2, 1, 1), b)
        self.layer2 = _tbr3d(128, 64, 4, 1, b)
        self.layer3 = _tbr3d(64, 32, 4, 1, b)
        self.layer4 = _tbr3d(32, 8, 4, 1, b)
        self.layer5 = nn.Sequential(nn.ConvTranspose3d(8, 1, 1, bias=b))

    def forward(self, feats):
        B, V = feats.shape[:2]
        g = self.spatial_reduce(feats.reshape(B * V, *feats.shape[2:]))
        g = g.unsqueeze(2).expand(-1, -1, 2, -1, -1).contiguous()
        raw = self.layer4(self.layer3(self.layer2(self.layer1(g))))
        gen = self.layer5(raw)
        raw = torch.cat((raw, gen), 1)
        return raw.reshape(B, V, 9, 32, 32, 32), gen.reshape(B, V, 32, 32, 32)


class RefMerger(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        blk = lambda cin, cout: nn.Sequential(nn.Conv3d(cin, cout, 3, padding=1), nn.BatchNorm3d(cout),
                                              nn.LeakyReLU(cfg.NETWORK.LEAKY_VALUE))
        self.layer1, self.layer2, self.layer3, self.layer4 = blk(9, 9), blk(9, 9), blk(9, 9), blk(9, 9)
        self.layer5, self.layer6 = blk(36, 9), blk(9, 1)

    def forward(self, raw, coarse, taps=None):
        B, V = raw.shape[:2]
        w1 = self.layer1(raw.reshape(B * V, 9, 32, 32, 32))
        w2 = self.layer2(w1)
        w3 = self.layer3(w2)
        w4 = self.layer4(w3)
        w = self.layer6(self.layer5(torch.cat((w1, w2, w3, w4), 1))).reshape(B, V, 32, 32, 32)
        if taps is not None:
            taps.update(merger_weights=w)
        return (coarse * torch.softmax(w, dim=1)).sum(1)


class RefRefiner(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        lv, b = cfg.NETWORK.LEAKY_VALUE, cfg.NETWORK.TCONV_USE_BIAS
        down = lambda cin, cout: nn.Sequential(nn.Conv3d(cin, cout, 4, padding=2), nn.BatchNorm3d(cout),
                                               nn.LeakyReLU(lv), nn.MaxPool3d(2))
        self.layer1, self.layer2, self.layer3 = down(1, 32), down(32, 64), down(64, 128)
        self.layer4 = nn.Sequential(nn.Linear(8192, 2048), nn.ReLU())
        self.layer5 = nn.Sequential(nn.Linear(2048, 8192), nn.ReLU())
        self.layer6 = _tbr3d(128, 64, 4, 1, b)
        self.layer7 = _tbr3d(64, 32, 4, 1, b)
        self.layer8 = nn.Sequential(nn.ConvTranspose3d(32, 1, 4, 2, 1, bias=b))

    def forward(self, vol, taps=None):
        l32 = vol.unsqueeze(1)
        l16 = self.layer1(l32)
        l8 = self.layer2(l16)
        l4 = self.layer3(l8)
        fc = self.layer5(self.layer4(l4.reshape(-1, 8192)))
        r4 = l4 + fc.reshape(-1, 128, 4, 4, 4)
        r8 = l8 + self.layer6(r4)
        r16 = l16 + self.layer7(r8)
        if taps is not None:
            taps.update(ref_l16=l16, ref_l8=l8, ref_l4=l4, ref_r4=r4, ref_r8=r8, ref_r16=r16)
        return ((l32 + self.layer8(r16)) * 0.5).squeeze(1)


def voxel_metrics(logits, gt, thresholds=(0.2, 0.3, 0.4, 0.5)):
    """core/test.py:141-164 per object.  Returns (counts int64 [B,T,5] = I,U,TP,FP,FN; iou [B,T]; f1 [B,T])."""
    prob = torch.sigmoid(logits).flatten(1)
    gt = gt.flatten(1).float()
    counts, ious, f1s = [], [], []
    for th in thresholds:
        v = torch.ge(prob, th).float()
        inter = (v * gt).sum(1)
        union = torch.ge(v + gt, 1).float().sum(1)
        iou = torch.where(union > 0, inter / union.clamp_min(1), torch.where(inter == 0, torch.ones_like(inter),
                                                                             torch.zeros_like(inter)))
        tp, fp, fn = (v * gt).sum(1), (v * (1 - gt)).sum(1), ((1 - v) * gt).sum(1)
        prec, rec = tp / (tp + fp + 1e-8), tp / (tp + fn + 1e-8)
        f1s.append(2 * prec * rec / (prec + rec + 1e-8))
        ious.append(iou)
        counts.append(torch.stack((inter, union, tp, fp, fn), -1))
    return torch.stack(counts, 1).long(), torch.stack(ious, 1), torch.stack(f1s, 1)


def forward_pipeline(enc, dec, mer, ref, images, cfg):
    """core/test.py:120-130 (epoch gates are 0 in config.py:107-108)"""
    raw, gen = dec(enc(images))
    vol = mer(raw, gen) if cfg.NETWORK.USE_MERGER else gen.mean(1)
    if cfg.NETWORK.USE_REFINER:
        vol = ref(vol)
    return vol


def init_weights(m):
    """utils/helpers.py:20-44 restated (that file imports matplotlib, which is not installed)"""
    if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="leaky_relu", a=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
        m.weight.data *= 0.1
    elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Linear):
        nn.init.normal_(m.weight, 0, 0.01)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
        m.weight.data *= 0.1
