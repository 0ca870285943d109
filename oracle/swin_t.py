"""ORACLE (test infrastructure, not product code): Swin-T feature extractor restated from timm.

The reference builds its Swin backbone with
    timm.create_model('swin_tiny_patch4_window7_224', pretrained, features_only=True, out_indices=...)
(/root/reference/models/swin_transformer.py:19-24).  timm is an un-vendored, un-pinned dependency
(requirements.txt:9; the only pin is timm 1.0.15 in Notebooks/SwinVox.ipynb cell 45) and is not
installed here, so its published algorithm is restated below with timm's module/key layout
(patch_embed.{proj,norm}, layers_{s}.downsample.{norm,reduction}, layers_{s}.blocks.{j}.{norm1,attn.
{relative_position_bias_table,qkv,proj},norm2,mlp.{fc1,fc2}}) as dumped in notebook cell 68.

Parity of this restatement is pinned against torchvision.models.swin_t (an independent implementation
of the same architecture that IS installed) in tests/test_oracle.py; timm itself cannot be run here
("parity unpinned" at the timm boundary).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

WINDOW = 7
DEPTHS = (2, 2, 6, 2)
HEADS = (3, 6, 12, 24)
EMBED = 96


def relative_position_index(ws=WINDOW):
    """(dh + ws-1) * (2ws-1) + (dw + ws-1) for every (query, key) pair of a ws x ws window"""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = coords[:, :, None] - coords[:, None, :]
    return (rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)


def shift_mask(H, W, ws, shift):
    """0 / -100 mask [nW, ws*ws, ws*ws] of SW-MSA, labelled on the rolled map"""
    img = torch.zeros(H, W)
    label = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[hs, wsl] = label
            label += 1
    win = img.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class PatchEmbed(nn.Module):
    def __init__(self, in_chans=3, dim=EMBED):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, 4, 4)
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(self.proj(x).permute(0, 2, 3, 1))  # NHWC


class WindowAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * WINDOW - 1) ** 2, heads))
        self.register_buffer("relative_position_index", relative_position_index(), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def bias(self):
        n = WINDOW * WINDOW
        return self.relative_position_bias_table[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1)

    def forward(self, x, mask=None):  # x: [B_, 49, C]
        B_, n, Cc = x.shape
        qkv = self.qkv(x).reshape(B_, n, 3, self.heads, Cc // self.heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = q @ k.transpose(-2, -1) + self.bias().unsqueeze(0)
        if mask is not None:
            nW = mask.shape[0]
            attn = (attn.view(-1, nW, self.heads, n, n) + mask[None, :, None]).view(-1, self.heads, n, n)
        attn = attn.softmax(-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B_, n, Cc))


class Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, res, shift):
        super().__init__()
        self.res = res
        self.shift = shift if res > WINDOW else 0  # a 7x7 map is a single window: never shifted
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim)
        mask = shift_mask(res, res, WINDOW, self.shift) if self.shift else None
        self.register_buffer("attn_mask", mask, persistent=False)

    def forward(self, x):  # NHWC
        B, H, W, Cc = x.shape
        y = self.norm1(x)
        if self.shift:
            y = torch.roll(y, (-self.shift, -self.shift), (1, 2))
        y = y.view(B, H // WINDOW, WINDOW, W // WINDOW, WINDOW, Cc).permute(0, 1, 3, 2, 4, 5).reshape(-1, 49, Cc)
        y = self.attn(y, self.attn_mask)
        y = y.view(B, H // WINDOW, W // WINDOW, WINDOW, WINDOW, Cc).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, Cc)
        if self.shift:
            y = torch.roll(y, (self.shift, self.shift), (1, 2))
        x = x + y
        return x + self.mlp(self.norm2(x))


class PatchMerging(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(4 * dim)
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)

    def forward(self, x):
        B, H, W, Cc = x.shape
        x = x.reshape(B, H // 2, 2, W // 2, 2, Cc).permute(0, 1, 3, 4, 2, 5).flatten(3)
        return self.reduction(self.norm(x))


class Stage(nn.Module):
    def __init__(self, s):
        super().__init__()
        dim, res = EMBED * 2 ** s, 56 // 2 ** s
        self.downsample = PatchMerging(dim // 2) if s > 0 else nn.Identity()
        self.blocks = nn.Sequential(*[Block(dim, HEADS[s], res, 0 if j % 2 == 0 else WINDOW // 2)
                                      for j in range(DEPTHS[s])])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class FeatureInfo:
    def __init__(self, chans):
        self._chans = chans

    def channels(self):
        return list(self._chans)


class SwinTFeatures(nn.Module):
    """timm FeatureListNet(flatten_sequential=True) view of swin_tiny_patch4_window7_224: children
    patch_embed, layers_0..; modules after the last requested stage are dropped; outputs are NHWC."""

    def __init__(self, out_indices=(0, 1, 2, 3)):
        super().__init__()
        self.out_indices = [i % 4 for i in out_indices]
        self.patch_embed = PatchEmbed()
        for s in range(max(self.out_indices) + 1):
            setattr(self, f"layers_{s}", Stage(s))
        self.feature_info = FeatureInfo([EMBED * 2 ** i for i in self.out_indices])
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        assert x.shape[-2:] == (224, 224), "swin_tiny_patch4_window7_224 is strict about its input size"
        x = self.patch_embed(x)
        feats = {}
        for s in range(max(self.out_indices) + 1):
            x = getattr(self, f"layers_{s}")(x)
            feats[s] = x
        return [feats[i] for i in self.out_indices]


def create_model(name, pretrained=False, features_only=True, out_indices=(0, 1, 2, 3), **_):
    """Drop-in for the single timm call the reference makes.  `pretrained` cannot be honoured offline."""
    assert name == "swin_tiny_patch4_window7_224" and features_only
    return SwinTFeatures(out_indices)


def load_from_torchvision(model, tv):
    """Copy torchvision.models.swin_t weights into the timm-layout module (key remap of SURVEY 8c)."""
    sd = {}
    tsd = tv.state_dict()

    def cp(dst, src):
        for suffix in ("weight", "bias"):
            if f"{src}.{suffix}" in tsd:
                sd[f"{dst}.{suffix}"] = tsd[f"{src}.{suffix}"]

    cp("patch_embed.proj", "features.0.0")
    cp("patch_embed.norm", "features.0.2")
    for s in range(4):
        if not hasattr(model, f"layers_{s}"):
            break
        if s > 0:
            cp(f"layers_{s}.downsample.norm", f"features.{2 * s}.norm")
            cp(f"layers_{s}.downsample.reduction", f"features.{2 * s}.reduction")
        for j in range(DEPTHS[s]):
            a, b = f"layers_{s}.blocks.{j}", f"features.{2 * s + 1}.{j}"
            for n in ("norm1", "norm2", "attn.qkv", "attn.proj"):
                cp(f"{a}.{n}", f"{b}.{n}")
            sd[f"{a}.attn.relative_position_bias_table"] = tsd[f"{b}.attn.relative_position_bias_table"]
            cp(f"{a}.mlp.fc1", f"{b}.mlp.0")
            cp(f"{a}.mlp.fc2", f"{b}.mlp.3")
    missing, unexpected = model.load_state_dict(sd, strict=True)
    return model
