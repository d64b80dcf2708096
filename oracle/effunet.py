"""Oracle (test infrastructure): restatement of smp.Unet on a timm EfficientNet.

Restates ``segmentation_models_pytorch==0.5.0`` ``Unet(encoder_name=
"timm-efficientnet-b{0,1,3,7}", classes=1, encoder_weights=None)`` built on
``timm==1.0.19`` (pins: reference ``uv.lock:1496-1497`` / ``uv.lock:1683-1684``).
Call site replaced: reference
``src/human_edge_detection/advanced/hierarchical_segmentation_unet.py:1770-1774``.

Neither package is installed in this image -> PARITY UNPINNED against the
packages themselves.  Evidence used instead:
  * state-dict key names/counts stated by the reference
    (``hierarchical_segmentation_unet.py:1815-1828``: B0~358, B1~506, B3~572,
    B7~1198 encoder keys; ``export_peopleseg_onnx.py:111-136``: decoder key
    pattern ``decoder.blocks.{i}.conv1.0.weight`` and channels 256..16);
  * numerical equality of the encoder with ``torchvision.models.efficientnet_b*``
    after a key remap (tests/test_oracle_effunet.py).

Published structure restated (timm ``efficientnet.py``/``_efficientnet_builder.py``
/``_efficientnet_blocks.py`` and smp ``encoders/timm_efficientnet.py``,
``decoders/unet/{model,decoder}.py``, ``base/{modules,heads}.py``):
  * stem ``conv_stem`` 3x3 s2 (no bias) + ``bn1`` + SiLU, stem width
    round_channels(32*w);
  * 7 stages, arch ``ds_r1_k3_s1_e1_c16 / ir_r2_k3_s2_e6_c24 / ir_r2_k5_s2_e6_c40 /
    ir_r3_k3_s2_e6_c80 / ir_r3_k5_s1_e6_c112 / ir_r4_k5_s2_e6_c192 /
    ir_r1_k3_s1_e6_c320`` with se_ratio 0.25 taken on the *block input* width,
    repeats ceil(r*depth_mult), widths round_channels(c*width_mult, 8);
  * symmetric padding ((s-1)+(k-1))//2, BatchNorm eps 1e-5, drop-path identity in eval;
  * features after stem and after stages {0-1},{2},{3-4},{5-6};
  * decoder: 5 blocks (256,128,64,32,16): nearest resize to the skip's HxW,
    concat skip, 2x[conv3x3 no-bias + BN + ReLU]; head conv3x3 16->1 with bias.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

_ARCH = [  # (kind, repeats, kernel, stride, expand, out_ch)
    ("ds", 1, 3, 1, 1, 16),
    ("ir", 2, 3, 2, 6, 24),
    ("ir", 2, 5, 2, 6, 40),
    ("ir", 3, 3, 2, 6, 80),
    ("ir", 3, 5, 1, 6, 112),
    ("ir", 4, 5, 2, 6, 192),
    ("ir", 1, 3, 1, 6, 320),
]
_SCALING = {"b0": (1.0, 1.0), "b1": (1.0, 1.1), "b2": (1.1, 1.2), "b3": (1.2, 1.4),
            "b4": (1.4, 1.8), "b5": (1.6, 2.2), "b6": (1.8, 2.6), "b7": (2.0, 3.1)}
DECODER_CHANNELS = (256, 128, 64, 32, 16)


def round_channels(c: float, mult: float = 1.0, div: int = 8, limit: float = 0.9) -> int:
    """timm ``round_channels``/``make_divisible`` (round_limit 0.9)."""
    if not mult:
        return int(c)
    v = c * mult
    new_v = max(div, int(v + div / 2) // div * div)
    if new_v < limit * v:
        new_v += div
    return new_v


def variant_of(encoder_name: str) -> str:
    name = encoder_name.lower()
    for k in _SCALING:
        if name.endswith(k):
            return k
    raise ValueError(f"unsupported encoder {encoder_name!r}")


def block_plan(variant: str):
    """Returns (stem_ch, [[(kind,k,s,exp,cin,cout), ...] per stage])."""
    w, d = _SCALING[variant]
    stem = round_channels(32, w)
    cin = stem
    stages = []
    for kind, r, k, s, e, c in _ARCH:
        cout = round_channels(c, w)
        blocks = []
        for i in range(int(math.ceil(r * d))):
            blocks.append((kind, k, s if i == 0 else 1, e, cin, cout))
            cin = cout
        stages.append(blocks)
    return stem, stages


class _BNAct(nn.BatchNorm2d):
    """timm BatchNormAct2d: BN followed by an optional activation (same keys as BN)."""

    def __init__(self, c, act=True):
        super().__init__(c, eps=1e-5)
        self.apply_act = act

    def forward(self, x):
        x = super().forward(x)
        return F.silu(x) if self.apply_act else x


class _SE(nn.Module):
    def __init__(self, c, rd):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, rd, 1)
        self.conv_expand = nn.Conv2d(rd, c, 1)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(F.silu(self.conv_reduce(s)))
        return x * torch.sigmoid(s)


class _DS(nn.Module):
    def __init__(self, k, s, cin, cout):
        super().__init__()
        self.has_skip = s == 1 and cin == cout
        self.conv_dw = nn.Conv2d(cin, cin, k, s, ((s - 1) + (k - 1)) // 2, groups=cin, bias=False)
        self.bn1 = _BNAct(cin)
        self.se = _SE(cin, round(cin * 0.25))
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = _BNAct(cout, act=False)

    def forward(self, x):
        y = self.bn2(self.conv_pw(self.se(self.bn1(self.conv_dw(x)))))
        return y + x if self.has_skip else y


class _IR(nn.Module):
    def __init__(self, k, s, e, cin, cout):
        super().__init__()
        mid = round_channels(cin * e)
        self.has_skip = s == 1 and cin == cout
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = _BNAct(mid)
        self.conv_dw = nn.Conv2d(mid, mid, k, s, ((s - 1) + (k - 1)) // 2, groups=mid, bias=False)
        self.bn2 = _BNAct(mid)
        # se_ratio is relative to the block input: rd = round(mid * 0.25 / e)
        self.se = _SE(mid, round(mid * (0.25 / e)))
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = _BNAct(cout, act=False)

    def forward(self, x):
        y = self.bn1(self.conv_pw(x))
        y = self.se(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(y))
        return y + x if self.has_skip else y


class Encoder(nn.Module):
    def __init__(self, variant: str):
        super().__init__()
        stem, stages = block_plan(variant)
        w, _ = _SCALING[variant]
        self.conv_stem = nn.Conv2d(3, stem, 3, 2, 1, bias=False)
        self.bn1 = _BNAct(stem)
        self.blocks = nn.Sequential(*[
            nn.Sequential(*[(_DS(k, s, ci, co) if kind == "ds" else _IR(k, s, e, ci, co))
                            for kind, k, s, e, ci, co in st]) for st in stages])
        last = stages[-1][-1][5]
        # present in the state dict, unused in forward (classifier is deleted by smp)
        self.conv_head = nn.Conv2d(last, round_channels(1280, w), 1, bias=False)
        self.bn2 = _BNAct(round_channels(1280, w))
        self.out_channels = (3, stem, stages[1][-1][5], stages[2][-1][5], stages[4][-1][5], last)

    def forward(self, x) -> List[torch.Tensor]:
        feats = [x]
        x = self.bn1(self.conv_stem(x)); feats.append(x)
        for idx, stage in enumerate(self.blocks):
            x = stage(x)
            if idx in (1, 2, 4, 6):
                feats.append(x)
        return feats


class _ConvBNReLU(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _ConvBNReLU(cin + cskip, cout)
        self.conv2 = _ConvBNReLU(cout, cout)

    def forward(self, x, size, skip=None):
        x = F.interpolate(x, size=size, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], 1)
        return self.conv2(self.conv1(x))


class Decoder(nn.Module):
    def __init__(self, enc_channels):
        super().__init__()
        enc = list(enc_channels[1:])[::-1]
        ins = [enc[0]] + list(DECODER_CHANNELS[:-1])
        skips = enc[1:] + [0]
        self.blocks = nn.ModuleList([_DecBlock(i, s, o) for i, s, o in zip(ins, skips, DECODER_CHANNELS)])

    def forward(self, feats):
        sizes = [f.shape[2:] for f in feats][::-1]
        feats = feats[1:][::-1]
        x, skips = feats[0], feats[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, tuple(sizes[i + 1]), skips[i] if i < len(skips) else None)
        return x


class Unet(nn.Module):
    """Drop-in for ``smp.Unet(encoder_name, classes, encoder_weights=None)`` (same keys)."""

    def __init__(self, encoder_name="timm-efficientnet-b3", classes=1, encoder_weights=None, **_):
        super().__init__()
        self.encoder = Encoder(variant_of(encoder_name))
        self.decoder = Decoder(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(DECODER_CHANNELS[-1], classes, 3, padding=1),
                                               nn.Identity(), nn.Identity())

    def forward(self, x):
        return self.segmentation_head(self.decoder(self.encoder(x)))
