"""Parameter containers that reproduce the reference's state-dict key names and shapes.

The modules here only *hold* parameters (``nn.Conv2d`` / ``nn.BatchNorm2d`` are used as typed
containers so shapes, dtypes and default initialisation match the reference); their
``forward`` is never called -- the compute runs in ``libhis_b200.so`` (see ``engine.py``).

Key-name sources in the reference (``hed/`` = src/human_edge_detection/):
  * hed/advanced/hierarchical_segmentation_rgb.py:564-727 (model), :657-673 (feature extractor)
  * hed/advanced/hierarchical_segmentation_refinement.py:31-55, 255-344, 434-548, 609-732
  * hed/advanced/hierarchical_segmentation_unet.py:35-58, 277-372, 1708-1971
  * smp 0.5.0 / timm 1.0.19 EfficientNet-UNet key scheme (evidence: ..._unet.py:1815-1828,
    export_peopleseg_onnx.py:111-136)
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn


class Slot(nn.Identity):
    """Index placeholder for parameter-free reference modules (activations, dropout, sigmoid)."""


class LayerNorm2dParams(nn.Module):
    """hed/model.py:18-38 -- affine kept as [1,C,1,1]."""

    def __init__(self, c: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(1, c, 1, 1))
        self.bias = nn.Parameter(torch.zeros(1, c, 1, 1))
        self.eps = 1e-5


class SpatialGroupNormParams(nn.Module):
    """SpatialGroupNorm (normalization_comparison.py:54-74): nn.GroupNorm under ``.norm``; asserts divisibility like the reference."""

    def __init__(self, c: int, groups: int = 8):
        super().__init__()
        assert c % groups == 0, f"num_channels ({c}) must be divisible by num_groups ({groups})"
        self.norm = nn.GroupNorm(groups, c)


class MixedNormParams(nn.Module):
    """MixedNormalization (normalization_comparison.py:129-147): eval mode returns batch_norm(x) alone (:143-147); the
    instance-norm affine exists only as parameters."""

    def __init__(self, c: int):
        super().__init__()
        self.batch_norm = nn.BatchNorm2d(c)
        self.instance_norm = nn.InstanceNorm2d(c, affine=True)


class AdaptiveInstanceNormParams(nn.Module):
    """AdaptiveInstanceNorm2d (normalization_comparison.py:12-57): per-(sample, channel) statistics + affine; the running
    statistics are buffers the forward never reads (they exist in the state dict)."""

    def __init__(self, c: int, eps: float = 1e-5):
        super().__init__()
        self.num_features, self.eps = c, eps
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))


class ForegroundAwareNormParams(nn.Module):
    """ForegroundAwareNorm (normalization_comparison.py:84-132): instance norm (no affine) whose per-channel scale / bias are blended
    per PIXEL between a foreground and a background set by a learned detector on the un-normalised input:
    p = sigmoid(conv1x1(relu(conv1x1(x)))),  y = IN(x) * (p*fg_scale + (1-p)*bg_scale) + (p*fg_bias + (1-p)*bg_bias)."""

    def __init__(self, c: int, eps: float = 1e-5):
        super().__init__()
        self.num_features, self.eps = c, eps
        self.norm = nn.InstanceNorm2d(c, eps=eps, affine=False)
        self.fg_scale = nn.Parameter(torch.ones(c))
        self.fg_bias = nn.Parameter(torch.zeros(c))
        self.bg_scale = nn.Parameter(torch.ones(c))
        self.bg_bias = nn.Parameter(torch.zeros(c))
        self.fg_detector = nn.Sequential(nn.Conv2d(c, c // 4, 1), nn.ReLU(inplace=True), nn.Conv2d(c // 4, 1, 1), nn.Sigmoid())


def norm_spec(kind: str, groups: int = 8) -> str:
    """Internal spelling of (normalization_type, normalization_groups) handed down the parameter tree: 'group:8'."""
    return f"{kind}:{int(groups)}"


def norm_params(kind: str, c: int, clamp: bool = True) -> nn.Module:
    """hed/advanced/normalization_comparison.py:159-206.  ``kind`` may carry the group count ('group:8', see norm_spec).
    clamp: the call site passes ``num_groups=min(normalization_groups, c)`` (most do); ResidualBlock, the contour / distance
    branches, shared_features and RGBFeatureExtractor pass normalization_groups as is."""
    k, _, gs = kind.lower().partition(":")
    groups = int(gs) if gs else 8
    if clamp:
        groups = min(groups, c)
    if k in ("layer", "layernorm", "layernorm2d"):
        return LayerNorm2dParams(c)
    if k in ("batch", "batchnorm", "batchnorm2d"):
        return nn.BatchNorm2d(c)
    if k in ("instance", "instancenorm", "instancenorm2d"):
        # one group per channel (his_groupnorm_act).  Runs in the strict precision mode only: with single-fp16 activations a channel
        # whose deviation is small against its mean loses its signal to the 2^-11 rounding of the pre-norm tensor before the
        # statistics are taken (measured 1.5e-2 .. 4e-2 against the reference); the split-fp16 pre-norm tensor carries ~21 bits.
        return nn.InstanceNorm2d(c, affine=True)
    if k == "adaptive_instance":
        return AdaptiveInstanceNormParams(c)
    if k in ("group", "groupnorm"):
        if c % groups != 0:                      # :188-193
            for g in (8, 4, 2, 1):
                if c % g == 0:
                    groups = g
                    break
        return nn.GroupNorm(groups, c)
    if k == "spatial_group":
        return SpatialGroupNormParams(c, groups)
    if k == "mixed":
        return MixedNormParams(c)
    if k == "foreground_aware":
        return ForegroundAwareNormParams(c)
    raise ValueError(f"Unknown normalization type: {k}")


def group_norm_args(norm: nn.Module):
    """(groups, weight[C], bias[C], eps) of a per-sample statistic norm, or None."""
    if isinstance(norm, SpatialGroupNormParams):
        norm = norm.norm
    if isinstance(norm, nn.GroupNorm):
        return norm.num_groups, norm.weight, norm.bias, norm.eps
    if isinstance(norm, (nn.InstanceNorm2d, AdaptiveInstanceNormParams)):       # one group per channel
        return norm.num_features, norm.weight, norm.bias, norm.eps
    return None


INSTANCE_NORMS = (nn.InstanceNorm2d, AdaptiveInstanceNormParams, ForegroundAwareNormParams)
NORM_MODULES = (nn.BatchNorm2d, LayerNorm2dParams, nn.GroupNorm, SpatialGroupNormParams, MixedNormParams) + INSTANCE_NORMS


def check_activation(name: str) -> str:
    n = name.lower()
    if n not in ("relu", "swish", "gelu", "silu"):
        raise ValueError(f"Unknown activation function: {n}")
    return n


class ResidualBlockParams(nn.Module):
    def __init__(self, c: int, norm: str):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1)
        self.norm1 = norm_params(norm, c, clamp=False)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1)
        self.norm2 = norm_params(norm, c, clamp=False)


class EnhancedUNetParams(nn.Module):
    """..._unet.py:277-372."""

    def __init__(self, cin: int, base: int, depth: int, norm: str):
        super().__init__()
        self.depth = depth
        ch = [cin] + [base * (2 ** i) for i in range(depth)]
        self.channels = ch
        self.encoders = nn.ModuleList()
        for i in range(depth):
            if i == 0:
                enc = nn.Sequential(nn.Conv2d(ch[0], ch[1], 3, padding=1), norm_params(norm, ch[1]), Slot(),
                                    ResidualBlockParams(ch[1], norm), ResidualBlockParams(ch[1], norm))
            else:
                enc = nn.Sequential(ResidualBlockParams(ch[i], norm), ResidualBlockParams(ch[i], norm),
                                    nn.Conv2d(ch[i], ch[i + 1], 3, padding=1), norm_params(norm, ch[i + 1]), Slot())
            self.encoders.append(enc)
        self.pools = nn.ModuleList([Slot() for _ in range(depth - 1)])
        self.bottleneck = nn.Sequential(ResidualBlockParams(ch[-1], norm), ResidualBlockParams(ch[-1], norm),
                                        nn.Conv2d(ch[-1], ch[-1], 3, padding=1), norm_params(norm, ch[-1]), Slot(),
                                        nn.Conv2d(ch[-1], ch[-1], 1), Slot())
        self.bottleneck_conv = nn.Conv2d(ch[-1], ch[-1], 3, padding=1)
        self.upconvs = nn.ModuleList()
        self.decoders = nn.ModuleList()
        for i in range(depth - 1, 0, -1):
            self.upconvs.append(nn.ConvTranspose2d(ch[i + 1], ch[i], 2, stride=2))
            self.decoders.append(nn.Sequential(nn.Conv2d(ch[i] * 2, ch[i], 3, padding=1), norm_params(norm, ch[i]), Slot(),
                                               ResidualBlockParams(ch[i], norm), ResidualBlockParams(ch[i], norm)))
        self.final = nn.Sequential(nn.Conv2d(ch[1], ch[1] // 2, 3, padding=1), norm_params(norm, ch[1] // 2), Slot(),
                                   nn.Conv2d(ch[1] // 2, 2, 1))


class ChannelAttentionParams(nn.Module):
    def __init__(self, c: int, reduction: int = 8, min_channels: int = 8):
        super().__init__()
        r = max(c // reduction, min_channels)
        self.fc1 = nn.Conv2d(c, r, 1, bias=False)
        self.activation = Slot()
        self.fc2 = nn.Conv2d(r, c, 1, bias=False)
        self.sigmoid = Slot()


class SpatialAttentionParams(nn.Module):
    def __init__(self, k: int = 7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, k, padding=k // 2, bias=False)
        self.sigmoid = Slot()


class BaseHeadParams(nn.Module):
    """ExtendedHierarchicalSegmentationHeadUNetV2, ..._refinement.py:434-548."""

    def __init__(self, cin: int, mid: int, norm: str, attention: bool, base: int, depth: int):
        super().__init__()
        self.shared_features = nn.Sequential(nn.Conv2d(cin, mid, 3, padding=1), norm_params(norm, mid), Slot(), Slot(),
                                             ResidualBlockParams(mid, norm), Slot(), ResidualBlockParams(mid, norm))
        self.bg_vs_fg_unet = EnhancedUNetParams(mid, base, depth, norm)
        self.upsample_bg_fg = nn.Sequential(nn.ConvTranspose2d(2, 32, 2, stride=2), norm_params(norm, 32), Slot(), nn.Conv2d(32, 2, 1))
        h = mid // 2
        if attention:
            self.target_vs_nontarget_branch = nn.ModuleList([
                ResidualBlockParams(mid, norm), SpatialAttentionParams(7), Slot(), nn.ConvTranspose2d(mid, h, 2, stride=2),
                norm_params(norm, h), Slot(), ChannelAttentionParams(h, 8), Slot(), ResidualBlockParams(h, norm), nn.Conv2d(h, 2, 1)])
        else:
            self.target_vs_nontarget_branch = nn.Sequential(
                ResidualBlockParams(mid, norm), Slot(), nn.ConvTranspose2d(mid, h, 2, stride=2), norm_params(norm, h), Slot(), Slot(),
                ResidualBlockParams(h, norm), nn.Conv2d(h, 2, 1))
        self.fg_gate = nn.Sequential(nn.Conv2d(2, mid // 4, 1), Slot(), Slot(), nn.Conv2d(mid // 4, mid // 2, 1), Slot(),
                                     nn.Conv2d(mid // 2, mid, 1), Slot())


class ContourBranchParams(nn.Module):
    def __init__(self, cin: int, c: int, norm: str):
        super().__init__()
        self.contour_branch = nn.Sequential(nn.Conv2d(cin, c, 3, padding=1), norm_params(norm, c), Slot(),
                                            nn.Conv2d(c, c, 3, padding=1), norm_params(norm, c), Slot(), nn.Conv2d(c, 1, 1), Slot())


class DistanceDecoderParams(nn.Module):
    def __init__(self, cin: int, c: int, norm: str):
        super().__init__()
        self.distance_head = nn.Sequential(nn.Conv2d(cin, c, 3, padding=1), norm_params(norm, c), Slot(),
                                           ResidualBlockParams(c, norm), nn.Conv2d(c, 1, 1))
        self.threshold = nn.Parameter(torch.tensor(0.3))


class BoundaryRefinerParams(nn.Module):
    """BoundaryRefinementModule, ..._refinement.py:58-92."""

    def __init__(self, norm: str, c: int = 32):
        super().__init__()
        self.edge_conv = nn.Sequential(nn.Conv2d(3, c, 3, padding=1), norm_params(norm, c), Slot(), nn.Conv2d(c, c, 3, padding=1),
                                       norm_params(norm, c), Slot(), nn.Conv2d(c, 3, 1))
        self.blend_weight = nn.Parameter(torch.tensor(0.01))


class ProgressiveDecoderParams(nn.Module):
    """ProgressiveUpsamplingDecoder, ..._refinement.py:152-190."""

    def __init__(self, cin: int, norm: str, num_classes: int = 3):
        super().__init__()
        self.stages = nn.ModuleList([
            nn.Sequential(nn.ConvTranspose2d(cin, cin // 2, 4, stride=2, padding=1), norm_params(norm, cin // 2), Slot(), ResidualBlockParams(cin // 2, norm)),
            nn.Sequential(nn.ConvTranspose2d(cin // 2, cin // 4, 4, stride=2, padding=1), norm_params(norm, cin // 4), Slot(), ResidualBlockParams(cin // 4, norm)),
            nn.Conv2d(cin // 4, num_classes, 1)])


class SubPixelDecoderParams(nn.Module):
    """SubPixelDecoder, ..._refinement.py:218-240."""

    def __init__(self, cin: int, num_classes: int = 3, upscale: int = 2):
        super().__init__()
        self.conv = nn.Conv2d(cin, num_classes * upscale ** 2, 3, padding=1)


class RefinedHeadParams(nn.Module):
    """RefinedHierarchicalSegmentationHead, ..._refinement.py:609-732 (module order = the reference constructor's)."""

    def __init__(self, cin, mid, norm, attention, contour, distance, base, depth, boundary=False, subpixel=False, progressive=False):
        super().__init__()
        self.base_head = BaseHeadParams(cin, mid, norm, attention, base, depth)
        if boundary:
            self.boundary_refiner = BoundaryRefinerParams(norm)
        if progressive:
            self.progressive_decoder = ProgressiveDecoderParams(mid, norm)
        if subpixel:
            self.subpixel_decoder = SubPixelDecoderParams(mid)
        if contour:
            self.contour_branch = ContourBranchParams(mid, 64, norm)
        if distance:
            self.distance_decoder = DistanceDecoderParams(mid, 128, norm)


class RGBFeatureExtractorParams(nn.Module):
    """RGBFeatureExtractor, rgb.py:221-295 (num_layers=4: 3 -> 64 -> 128 -> 192 -> 256; a ResidualBlock after stages 1..3)."""

    def __init__(self, norm: str):
        super().__init__()
        ch = [3, 64, 128, 192, 256]
        layers = []
        for i in range(4):
            layers += [nn.Conv2d(ch[i], ch[i + 1], 3, padding=1), norm_params(norm, ch[i + 1]), Slot()]
            if i >= 1:
                layers.append(ResidualBlockParams(ch[i + 1], norm))
        self.features = nn.Sequential(*layers)


class FusionProjParams(nn.Sequential):
    """MultiScaleRGBSegmentationModel.fusion_proj, rgb.py:846-851."""

    def __init__(self, cin: int, cout: int, norm: str):
        super().__init__(nn.Conv2d(cin, cout, 1), norm_params(norm, cout), Slot())


class GuidedHeadParams(nn.Module):
    """PretrainedUNetGuidedSegmentationHead, rgb.py:43-123 (the head built when no refinement flag is set, :715-727)."""

    def __init__(self, cin: int, mid: int, norm: str, attention: bool):
        super().__init__()
        self.use_attention_module = attention
        self.input_adjust = nn.Conv2d(cin + 1, cin, 1)
        self.feature_processor = nn.Sequential(nn.Conv2d(cin, mid, 3, padding=1), norm_params(norm, mid), Slot(), Slot(),
                                               ResidualBlockParams(mid, norm), Slot(), ResidualBlockParams(mid, norm))
        self.final_classifier = nn.Sequential(nn.Conv2d(mid, mid // 2, 3, padding=1), norm_params(norm, mid // 2), Slot(),
                                              nn.Conv2d(mid // 2, 3, 1))
        if attention:
            self.attention_module = nn.Sequential(nn.Conv2d(mid, mid // 4, 1), Slot(), nn.Conv2d(mid // 4, 1, 1), Slot())
        with torch.no_grad():                      # rgb.py:117-123
            self.final_classifier[-1].bias.data[0] = 0.0
            self.final_classifier[-1].bias.data[1] = 0.0
            self.final_classifier[-1].bias.data[2] = -0.5


# ----------------------------------------------------------------------------- EfficientNet-UNet (smp/timm key scheme)
_ARCH = [("ds", 1, 3, 1, 1, 16), ("ir", 2, 3, 2, 6, 24), ("ir", 2, 5, 2, 6, 40), ("ir", 3, 3, 2, 6, 80),
         ("ir", 3, 5, 1, 6, 112), ("ir", 4, 5, 2, 6, 192), ("ir", 1, 3, 1, 6, 320)]
_SCALING = {"b0": (1.0, 1.0), "b1": (1.0, 1.1), "b2": (1.1, 1.2), "b3": (1.2, 1.4), "b4": (1.4, 1.8), "b5": (1.6, 2.2),
            "b6": (1.8, 2.6), "b7": (2.0, 3.1)}
DECODER_CHANNELS = (256, 128, 64, 32, 16)


def round_channels(c: float, mult: float = 1.0, div: int = 8) -> int:
    v = c * mult
    n = max(div, int(v + div / 2) // div * div)
    return n + div if n < 0.9 * v else n


def encoder_variant(encoder_name: str) -> str:
    n = encoder_name.lower()
    if "efficientnet" not in n:
        raise NotImplementedError(f"encoder {encoder_name!r}: only timm-efficientnet-b0..b7 are implemented")
    for k in _SCALING:
        if n.endswith(k):
            return k
    raise NotImplementedError(f"encoder {encoder_name!r}: only timm-efficientnet-b0..b7 are implemented")


class SEParams(nn.Module):
    def __init__(self, c, r):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, r, 1)
        self.conv_expand = nn.Conv2d(r, c, 1)


class DSBlockParams(nn.Module):
    kind = "ds"

    def __init__(self, k, s, cin, cout):
        super().__init__()
        self.k, self.s, self.cin, self.cout, self.mid = k, s, cin, cout, cin
        self.conv_dw = nn.Conv2d(cin, cin, k, s, ((s - 1) + (k - 1)) // 2, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin)
        self.se = SEParams(cin, round(cin * 0.25))
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)


class IRBlockParams(nn.Module):
    kind = "ir"

    def __init__(self, k, s, e, cin, cout):
        super().__init__()
        mid = round_channels(cin * e)
        self.k, self.s, self.cin, self.cout, self.mid = k, s, cin, cout, mid
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv_dw = nn.Conv2d(mid, mid, k, s, ((s - 1) + (k - 1)) // 2, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.se = SEParams(mid, round(mid * (0.25 / e)))
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)


class EncoderParams(nn.Module):
    def __init__(self, variant: str):
        super().__init__()
        w, d = _SCALING[variant]
        stem = round_channels(32, w)
        self.conv_stem = nn.Conv2d(3, stem, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(stem)
        stages, cin = [], stem
        for kind, r, k, s, e, c in _ARCH:
            cout = round_channels(c, w)
            blocks = []
            for i in range(int(math.ceil(r * d))):
                st = s if i == 0 else 1
                blocks.append(DSBlockParams(k, st, cin, cout) if kind == "ds" else IRBlockParams(k, st, e, cin, cout))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        self.conv_head = nn.Conv2d(cin, round_channels(1280, w), 1, bias=False)   # in the state dict, unused in forward
        self.bn2 = nn.BatchNorm2d(round_channels(1280, w))
        self.out_channels: Tuple[int, ...] = (3, stem, stages[1][-1].cout, stages[2][-1].cout, stages[4][-1].cout, stages[6][-1].cout)


class DecoderBlockParams(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.cin, self.cskip, self.cout = cin, cskip, cout
        self.conv1 = nn.Sequential(nn.Conv2d(cin + cskip, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), Slot())
        self.conv2 = nn.Sequential(nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), Slot())


class DecoderParams(nn.Module):
    def __init__(self, enc_channels):
        super().__init__()
        enc = list(enc_channels[1:])[::-1]
        ins = [enc[0]] + list(DECODER_CHANNELS[:-1])
        skips = enc[1:] + [0]
        self.blocks = nn.ModuleList([DecoderBlockParams(i, s, o) for i, s, o in zip(ins, skips, DECODER_CHANNELS)])


class SmpUnetParams(nn.Module):
    """Same keys as ``smp.Unet(encoder_name, classes=1, encoder_weights=None)`` (..._unet.py:1770-1774)."""

    def __init__(self, encoder_name: str):
        super().__init__()
        self.encoder = EncoderParams(encoder_variant(encoder_name))
        self.decoder = DecoderParams(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(DECODER_CHANNELS[-1], 1, 3, padding=1), Slot(), Slot())
