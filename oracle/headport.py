"""Oracle (test infrastructure): plain-torch fp32 restatement of the reference-owned hot path.

Functional style over a flat state dict (reference key names), so the same weights
drive the reference modules, this port and the CUDA model.  Pinned against the real
reference modules by tests/test_oracle_vs_reference.py (this container) and by the
golden fixtures in tests/golden/ (anywhere).  Citations are into /root/reference/
``src/human_edge_detection/`` (abbreviated ``hed/``).

Covered: the preset path (`HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet`
with the refined head), BatchNorm (eval) and LayerNorm2d normalisation, relu/silu/
swish/gelu activations, attention and non-attention target branches, contour and
distance branches, the export wrapper outputs and `MaskDilationModule`; and the
secondary variant the same factory reaches when no refinement flag is set
(`PretrainedUNetGuidedSegmentationHead`, rgb.py:43-218).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import effunet

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class PathConfig:
    """The kwargs of ``create_rgb_hierarchical_model`` that shape the preset path
    (hed/advanced/hierarchical_segmentation_rgb.py:925-1026, train_advanced.py:132-160)."""
    roi_size: Tuple[int, int] = (64, 48)
    mask_size: Tuple[int, int] = (128, 96)
    encoder_name: str = "timm-efficientnet-b0"
    pretrained_weights_path: str = "ext_extractor/best_model_b0_0.8741.pth"
    use_attention_module: bool = True
    use_contour_detection: bool = True
    use_distance_transform: bool = True
    normalization_type: str = "batchnorm"
    normalization_groups: int = 8
    activation_function: str = "relu"
    activation_beta: float = 1.0
    hierarchical_base_channels: int = 64
    hierarchical_depth: int = 3
    # DynamicRoIAlign.spatial_scale of both aligners: 640.0 by default (rgb.py:636-647);
    # the exporter overrides it with (H, W) (export_onnx_advanced.py:80-98).
    spatial_scale: Tuple[float, float] = (640.0, 640.0)
    # False -> HierarchicalRGBSegmentationModel (rgb.py:298-439): no UNet branch, one RoIAlign with aligned=False
    use_pretrained_unet: bool = True
    # True -> MultiScaleRGBSegmentationModel (rgb.py:777-922)
    use_boundary_refinement: bool = False      # BoundaryRefinementModule on the final logits (..._refinement.py:58-149)
    use_subpixel_conv: bool = False            # SubPixelDecoder re-decode of the shared features (:218-252)
    use_progressive_upsampling: bool = False   # ProgressiveUpsamplingDecoder re-decode (:152-215); wins over sub-pixel (:753-756)
    multi_scale: bool = False
    roi_sizes: tuple = (("scale1", 56), ("scale2", 42), ("scale3", 28))
    fusion_method: str = "concat"

    def factory_kwargs(self) -> dict:
        if self.multi_scale:
            return dict(mask_size=self.mask_size, multi_scale=True, roi_sizes=dict(self.roi_sizes), fusion_method=self.fusion_method,
                        use_attention_module=self.use_attention_module, normalization_type=self.normalization_type, normalization_groups=self.normalization_groups,
                        activation_function=self.activation_function, activation_beta=self.activation_beta)
        if not self.use_pretrained_unet:
            return dict(roi_size=self.roi_size, mask_size=self.mask_size, multi_scale=False,
                        use_attention_module=self.use_attention_module, use_contour_detection=self.use_contour_detection,
                        use_distance_transform=self.use_distance_transform, normalization_type=self.normalization_type,
                        normalization_groups=self.normalization_groups)
        return dict(roi_size=self.roi_size, mask_size=self.mask_size, multi_scale=False,
                    use_attention_module=self.use_attention_module,
                    use_boundary_refinement=self.use_boundary_refinement, use_subpixel_conv=self.use_subpixel_conv,
                    use_progressive_upsampling=self.use_progressive_upsampling,
                    use_contour_detection=self.use_contour_detection,
                    use_distance_transform=self.use_distance_transform,
                    normalization_type=self.normalization_type, normalization_groups=self.normalization_groups,
                    activation_function=self.activation_function, activation_beta=self.activation_beta,
                    use_pretrained_unet=True, pretrained_weights_path=self.pretrained_weights_path,
                    freeze_pretrained_weights=True, use_full_image_unet=True,
                    encoder_name=self.encoder_name,
                    hierarchical_base_channels=self.hierarchical_base_channels,
                    hierarchical_depth=self.hierarchical_depth)


PRESETS = {
    # hed/experiments/config_manager.py:2558-2589 (ModelConfig defaults :189-190 give bc=64, depth=3)
    "b0": PathConfig(),
    # config_manager.py:3643-3676
    "b1_enhanced": PathConfig(roi_size=(80, 60), mask_size=(160, 120), encoder_name="timm-efficientnet-b1",
                              pretrained_weights_path="ext_extractor/best_model_b1_0.8833.pth",
                              hierarchical_base_channels=72, hierarchical_depth=3),
    # config_manager.py:3851-3884
    "b7_ultra": PathConfig(roi_size=(128, 96), mask_size=(256, 192), encoder_name="timm-efficientnet-b7",
                           pretrained_weights_path="ext_extractor/best_model_b7_0.9009.pth",
                           hierarchical_base_channels=96, hierarchical_depth=4),
}


# ----------------------------------------------------------------------------- primitives
def act_rgb(x: Tensor, cfg: PathConfig) -> Tensor:
    """``get_activation_function`` (rgb.py:21-40, ..._unet.py:13-32): swish == silu, beta ignored."""
    a = cfg.activation_function.lower()
    if a == "relu":
        return F.relu(x)
    if a in ("swish", "silu"):
        return F.silu(x)
    if a == "gelu":
        return F.gelu(x)
    raise ValueError(f"Unsupported activation function: {a}")


def act_ref(x: Tensor, cfg: PathConfig) -> Tensor:
    """``get_activation`` (activation_utils.py:71-103): swish uses x*sigmoid(beta*x)."""
    a = cfg.activation_function.lower()
    if a == "relu":
        return F.relu(x)
    if a == "swish":
        return x * torch.sigmoid(cfg.activation_beta * x)
    if a == "gelu":
        return F.gelu(x)
    if a == "silu":
        return F.silu(x)
    raise ValueError(f"Unknown activation function: {a}")


def norm(sd: SD, p: str, x: Tensor, cfg: PathConfig) -> Tensor:
    """``get_normalization_layer`` (normalization_comparison.py:159-206), eval mode.
    batchnorm: nn.BatchNorm2d eps 1e-5.  layernorm2d: model.py:18-38 -- statistics over
    (C,H,W) per sample, biased variance, eps 1e-5, affine [1,C,1,1]."""
    t = cfg.normalization_type.lower()
    if t in ("batch", "batchnorm", "batchnorm2d"):
        return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                            False, 0.0, 1e-5)
    if t in ("layer", "layernorm", "layernorm2d"):
        mean = x.mean(dim=(1, 2, 3), keepdim=True)
        var = x.var(dim=(1, 2, 3), keepdim=True, unbiased=False)
        return (x - mean) / torch.sqrt(var + 1e-5) * sd[p + "weight"] + sd[p + "bias"]
    if t in ("instance", "instancenorm", "instancenorm2d"):
        # nn.InstanceNorm2d(affine=True) (:185-186): per (sample, channel) statistics over (H,W), biased variance, eps 1e-5
        return F.instance_norm(x, None, None, sd[p + "weight"], sd[p + "bias"], True, 0.0, 1e-5)
    if t == "adaptive_instance":
        # AdaptiveInstanceNorm2d.forward (:31-51), written out like the reference (mean / var of the flattened map, divide by
        # sqrt(var + eps)): same mathematics as instance norm, but the operation order matters at fp32 for near-constant
        # channels (F.instance_norm moves the small-ROI golden by 1.4e-3); the running statistics are never read
        b, c, h, w = x.shape
        xr = x.reshape(b, c, -1)
        mean = xr.mean(dim=2, keepdim=True)
        var = xr.var(dim=2, keepdim=True, unbiased=False)
        xn = ((xr - mean) / torch.sqrt(var + 1e-5)).view(b, c, h, w)
        return xn * sd[p + "weight"].view(1, c, 1, 1) + sd[p + "bias"].view(1, c, 1, 1)
    if t == "foreground_aware":
        # ForegroundAwareNorm.forward (:113-132): instance norm without affine, scale / bias blended per pixel by the fg detector
        c = x.shape[1]
        xn = F.instance_norm(x, None, None, None, None, True, 0.0, 1e-5)
        prob = torch.sigmoid(conv(sd, p + "fg_detector.2.", F.relu(conv(sd, p + "fg_detector.0.", x))))
        v = lambda k: sd[p + k].view(1, c, 1, 1)
        scale = prob * v("fg_scale") + (1 - prob) * v("bg_scale")
        bias = prob * v("fg_bias") + (1 - prob) * v("bg_bias")
        return xn * scale + bias
    if t in ("group", "groupnorm", "spatial_group"):
        # nn.GroupNorm (:187-194; SpatialGroupNorm :54-74 wraps one under .norm); group count as the factory resolves it for
        # group counts that do not exceed the channel count
        q = p + "norm." if t == "spatial_group" else p
        c, g = x.shape[1], min(cfg.normalization_groups, x.shape[1])
        if c % g and t != "spatial_group":
            g = next(d for d in (8, 4, 2, 1) if c % d == 0)
        return F.group_norm(x, g, sd[q + "weight"], sd[q + "bias"], 1e-5)
    if t == "mixed":     # MixedNormalization.forward in eval mode (:143-147) = its BatchNorm2d
        q = p + "batch_norm."
        return F.batch_norm(x, sd[q + "running_mean"], sd[q + "running_var"], sd[q + "weight"], sd[q + "bias"], False, 0.0, 1e-5)
    raise ValueError(f"oracle does not restate normalization type {t!r}")


def conv(sd: SD, p: str, x: Tensor, padding: int = 0) -> Tensor:
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), padding=padding)


def convT2(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.conv_transpose2d(x, sd[p + "weight"], sd.get(p + "bias"), stride=2)


def residual_block(sd: SD, p: str, x: Tensor, cfg: PathConfig, act) -> Tensor:
    """..._refinement.py:31-55 (act=act_ref) and ..._unet.py:35-58 (act=act_rgb): same dataflow."""
    y = act(norm(sd, p + "norm1.", conv(sd, p + "conv1.", x, 1), cfg), cfg)
    y = norm(sd, p + "norm2.", conv(sd, p + "conv2.", y, 1), cfg)
    return act(y + x, cfg)


def roi_align(feat: Tensor, rois: Tensor, oh: int, ow: int, scale_h: float, scale_w: float,
              aligned: bool = True) -> Tensor:
    """hed/dynamic_roi_align.py:56-171 restated without grid_sample/index_select.

    Output (i,j) of ROI k samples image ``batch_idx`` bilinearly (zero padding) at pixel
    ``fx = x1*Sw + j/(ow-1)*(x2-x1)*Sw``, ``fy = y1*Sh + i/(oh-1)*(y2-y1)*Sh`` -- an inclusive
    linspace grid, not bin centres.  The reference normalises to [-1,1] (:139-147) and
    grid_sample un-normalises again: aligned -> ((g+1)/2)*(size-1), else ((g+1)*size-1)/2.
    The round trip is kept so that float rounding matches.
    """
    K = rois.shape[0]
    B, C, H, W = feat.shape
    bidx = rois[:, 0].long()
    x1 = rois[:, 1] * scale_w
    y1 = rois[:, 2] * scale_h
    x2 = rois[:, 3] * scale_w
    y2 = rois[:, 4] * scale_h
    gx = torch.linspace(0, 1, ow)
    gy = torch.linspace(0, 1, oh)
    fx = x1[:, None] + gx[None, :] * (x2 - x1)[:, None]          # [K, ow]
    fy = y1[:, None] + gy[None, :] * (y2 - y1)[:, None]          # [K, oh]
    if aligned:
        nx = (fx / (W - 1)) * 2 - 1
        ny = (fy / (H - 1)) * 2 - 1
        px = ((nx + 1) / 2) * (W - 1)
        py = ((ny + 1) / 2) * (H - 1)
    else:
        nx = (fx / W) * 2 - 1
        ny = (fy / H) * 2 - 1
        px = ((nx + 1) * W - 1) / 2
        py = ((ny + 1) * H - 1) / 2
    x0 = torch.floor(px); y0 = torch.floor(py)
    wx1 = px - x0; wy1 = py - y0
    out = feat.new_zeros(K, C, oh, ow)
    if K == 0:
        return out
    src = feat[bidx]                                                # [K,C,H,W]
    flat = src.reshape(K, C, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = (x0 + dx).long(); yi = (y0 + dy).long()            # [K,ow], [K,oh]
            wx = wx1 if dx else 1 - wx1
            wy = wy1 if dy else 1 - wy1
            okx = (xi >= 0) & (xi < W); oky = (yi >= 0) & (yi < H)
            idx = yi.clamp(0, H - 1)[:, :, None] * W + xi.clamp(0, W - 1)[:, None, :]     # [K,oh,ow]
            w = (wy * oky)[:, :, None] * (wx * okx)[:, None, :]
            g = torch.gather(flat, 2, idx.reshape(K, 1, oh * ow).expand(K, C, oh * ow)).reshape(K, C, oh, ow)
            out += g * w[:, None]
    return out


# ----------------------------------------------------------------------------- UNet branch
def imagenet_or_half_norm(path: str):
    """..._unet.py:1744-1758: ImageNet mean/std iff the *weights path string* names b0/b1/b7."""
    if any(v in path.lower() for v in ("b0", "b1", "b7")):
        return [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    return [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]


def unet_state(sd: SD) -> SD:
    pre = "pretrained_unet.model.model."
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


_UNET_CACHE: dict = {}


def pretrained_unet_logits(sd: SD, images: Tensor, cfg: PathConfig) -> Tensor:
    """PreTrainedPeopleSegmentationUNetWrapper.forward (..._unet.py:1973-1992) ->
    PreTrainedPeopleSegmentationUNet.forward (:1901-1916): ``x/255 if x.max()>1``, (x-mean)/std,
    smp.Unet, then ``output_conv`` 1x1 (weights pinned to [+1,-1], bias 0 at :1963-1971).
    Returns the 2-channel logits [B,2,H,W]."""
    key = cfg.encoder_name
    net = _UNET_CACHE.get(key)
    if net is None:
        net = effunet.Unet(cfg.encoder_name, classes=1).eval()
        _UNET_CACHE[key] = net
    net.load_state_dict(unet_state(sd), strict=True)
    x = images
    if x.max() > 1.0:
        x = x / 255.0
    mean = sd.get("pretrained_unet.model.norm_mean")
    std = sd.get("pretrained_unet.model.norm_std")
    if mean is None:
        m, s = imagenet_or_half_norm(cfg.pretrained_weights_path)
        mean = torch.tensor(m).view(1, 3, 1, 1); std = torch.tensor(s).view(1, 3, 1, 1)
    with torch.no_grad():
        one = net((x - mean) / std)
    return F.conv2d(one, sd["pretrained_unet.output_conv.weight"], sd["pretrained_unet.output_conv.bias"])


# ----------------------------------------------------------------------------- head
def rgb_feature_extractor(sd: SD, x: Tensor, cfg: PathConfig) -> Tensor:
    """rgb.py:657-673: conv3x3 3->64,N,A,RB(64),conv3x3 64->128,N,A,RB(128),conv3x3 128->256,N,A,RB(256),
    conv1x1 256->256,N,A.  Stand-alone activations come from rgb.py's factory, the RBs are the
    refinement ones."""
    p = "rgb_feature_extractor."
    for i in (0, 4, 8):
        x = act_rgb(norm(sd, f"{p}{i + 1}.", conv(sd, f"{p}{i}.", x, 1), cfg), cfg)
        x = residual_block(sd, f"{p}{i + 3}.", x, cfg, act_ref)
    return act_rgb(norm(sd, p + "13.", conv(sd, p + "12.", x, 0), cfg), cfg)


def enhanced_unet(sd: SD, p: str, x: Tensor, cfg: PathConfig) -> Tensor:
    """..._unet.py:277-417 (ctor :313-372, forward :375-417)."""
    d = cfg.hierarchical_depth
    skips = []
    for i in range(d):
        e = f"{p}encoders.{i}."
        if i == 0:
            x = act_rgb(norm(sd, e + "1.", conv(sd, e + "0.", x, 1), cfg), cfg)
            x = residual_block(sd, e + "3.", x, cfg, act_rgb)
            x = residual_block(sd, e + "4.", x, cfg, act_rgb)
        else:
            x = residual_block(sd, e + "0.", x, cfg, act_rgb)
            x = residual_block(sd, e + "1.", x, cfg, act_rgb)
            x = act_rgb(norm(sd, e + "3.", conv(sd, e + "2.", x, 1), cfg), cfg)
        skips.append(x)
        if i < d - 1:
            x = F.max_pool2d(x, 2)
    b = p + "bottleneck."
    a = residual_block(sd, b + "0.", x, cfg, act_rgb)
    a = residual_block(sd, b + "1.", a, cfg, act_rgb)
    a = act_rgb(norm(sd, b + "3.", conv(sd, b + "2.", a, 1), cfg), cfg)
    a = torch.sigmoid(conv(sd, b + "5.", a, 0))
    x = conv(sd, p + "bottleneck_conv.", x, 1) * a
    for i in range(d - 1):
        x = convT2(sd, f"{p}upconvs.{i}.", x)
        skip = skips[d - 2 - i]
        x = F.interpolate(x, size=skip.shape[2:], mode="bilinear", align_corners=False)  # :408 (identity if equal)
        x = torch.cat([x, skip], 1)
        q = f"{p}decoders.{i}."
        x = act_rgb(norm(sd, q + "1.", conv(sd, q + "0.", x, 1), cfg), cfg)
        x = residual_block(sd, q + "3.", x, cfg, act_rgb)
        x = residual_block(sd, q + "4.", x, cfg, act_rgb)
    f = p + "final."
    x = act_rgb(norm(sd, f + "1.", conv(sd, f + "0.", x, 1), cfg), cfg)
    return conv(sd, f + "3.", x, 0)


def spatial_attention(sd: SD, p: str, x: Tensor) -> Tensor:
    """attention_modules.py:67-113: x * sigmoid(conv7x7([mean_c, max_c]))."""
    s = torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)
    return x * torch.sigmoid(F.conv2d(s, sd[p + "conv.weight"], None, padding=sd[p + "conv.weight"].shape[-1] // 2))


def channel_attention(sd: SD, p: str, x: Tensor, cfg: PathConfig) -> Tensor:
    """attention_modules.py:10-64: x * sigmoid(fc2(A(fc1(avgpool(x))))) (no biases)."""
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.conv2d(act_ref(F.conv2d(s, sd[p + "fc1.weight"]), cfg), sd[p + "fc2.weight"])
    return x * torch.sigmoid(s)


def _to_mask_size(x: Tensor, cfg: PathConfig) -> Tensor:
    mh, mw = cfg.mask_size
    if x.shape[2] != mh or x.shape[3] != mw:
        x = F.interpolate(x, size=(mh, mw), mode="bilinear", align_corners=False)
    return x


def base_head(sd: SD, p: str, feats: Tensor, cfg: PathConfig):
    """ExtendedHierarchicalSegmentationHeadUNetV2 (..._refinement.py:434-606; dropout = identity in eval)."""
    s = p + "shared_features."
    shared = act_ref(norm(sd, s + "1.", conv(sd, s + "0.", feats, 1), cfg), cfg)
    shared = residual_block(sd, s + "4.", shared, cfg, act_ref)
    shared = residual_block(sd, s + "6.", shared, cfg, act_ref)

    low = enhanced_unet(sd, p + "bg_vs_fg_unet.", shared, cfg)

    u = p + "upsample_bg_fg."
    bgfg = act_ref(norm(sd, u + "1.", convT2(sd, u + "0.", low), cfg), cfg)
    bgfg = _to_mask_size(conv(sd, u + "3.", bgfg, 0), cfg)
    probs = F.softmax(bgfg, dim=1)

    g = p + "fg_gate."
    gate = act_ref(conv(sd, g + "0.", low, 0), cfg)
    gate = act_ref(conv(sd, g + "3.", gate, 0), cfg)
    gate = torch.sigmoid(conv(sd, g + "5.", gate, 0))
    x = shared * gate

    t = p + "target_vs_nontarget_branch."
    if cfg.use_attention_module:      # :509-523
        x = residual_block(sd, t + "0.", x, cfg, act_ref)
        x = spatial_attention(sd, t + "1.", x)
        x = act_ref(norm(sd, t + "4.", convT2(sd, t + "3.", x), cfg), cfg)
        x = channel_attention(sd, t + "6.", x, cfg)
        x = residual_block(sd, t + "8.", x, cfg, act_ref)
        tn = conv(sd, t + "9.", x, 0)
    else:                             # :525-534
        x = residual_block(sd, t + "0.", x, cfg, act_ref)
        x = act_ref(norm(sd, t + "3.", convT2(sd, t + "2.", x), cfg), cfg)
        x = residual_block(sd, t + "6.", x, cfg, act_ref)
        tn = conv(sd, t + "7.", x, 0)
    tn = _to_mask_size(tn, cfg)

    fg = probs[:, 1]
    logits = torch.stack([bgfg[:, 0], bgfg[:, 1] + tn[:, 0] * fg, bgfg[:, 1] + tn[:, 1] * fg], 1)   # :588-596
    aux = {"bg_fg_logits": bgfg, "bg_fg_logits_low": low, "target_nontarget_logits": tn,
           "fg_attention": gate, "shared_features": shared}
    return logits, aux


def boundary_refine(sd: SD, p: str, logits: Tensor, cfg: PathConfig) -> Tensor:
    """BoundaryRefinementModule.forward (..._refinement.py:131-149): edge map from softmax gradients, normalised by the min / max
    over the WHOLE batch tensor (:118-128), gates a small conv net's correction."""
    probs = torch.softmax(logits, 1)
    dy = F.pad((probs[:, :, 1:] - probs[:, :, :-1]).abs(), (0, 0, 0, 1), mode="replicate")
    dx = F.pad((probs[:, :, :, 1:] - probs[:, :, :, :-1]).abs(), (0, 1, 0, 0), mode="replicate")
    edges = torch.sqrt(dy ** 2 + dx ** 2).mean(1, keepdim=True)
    emin, emax = edges.min(), edges.max()
    edges = torch.zeros_like(edges) if emax - emin < 1e-6 else (edges - emin) / (emax - emin + 1e-6)
    e = p + "edge_conv."
    x = act_ref(norm(sd, e + "1.", conv(sd, e + "0.", logits, 1), cfg), cfg)
    x = act_ref(norm(sd, e + "4.", conv(sd, e + "3.", x, 1), cfg), cfg)
    x = conv(sd, e + "6.", x, 0)
    return logits + sd[p + "blend_weight"] * x * edges


def refined_head(sd: SD, p: str, feats: Tensor, cfg: PathConfig, aux_branches: bool = True):
    """RefinedHierarchicalSegmentationHead.forward (..._refinement.py:734-804), preset flags only
    (no boundary refiner / progressive / sub-pixel decoders).  aux_branches=False skips the contour / distance branches, which
    feed aux outputs only: the exported ONNX graph (outputs: masks + binary masks, export_onnx_advanced.py:353-457) prunes them."""
    logits, aux = base_head(sd, p + "base_head.", feats, cfg)
    shared = aux["shared_features"]
    if cfg.use_progressive_upsampling:    # ProgressiveUpsamplingDecoder :152-215 replaces the hierarchical logits (:753-756)
        d = p + "progressive_decoder.stages."
        x = shared
        for i in (0, 1):
            x = F.conv_transpose2d(x, sd[f"{d}{i}.0.weight"], sd[f"{d}{i}.0.bias"], stride=2, padding=1)
            x = act_ref(norm(sd, f"{d}{i}.1.", x, cfg), cfg)
            x = residual_block(sd, f"{d}{i}.3.", x, cfg, act_ref)
        logits = _to_mask_size(conv(sd, d + "2.", x, 0), cfg)
    elif cfg.use_subpixel_conv:       # SubPixelDecoder :218-252 replaces the hierarchical logits (:757-763)
        logits = _to_mask_size(F.pixel_shuffle(conv(sd, p + "subpixel_decoder.conv.", shared, 1), 2), cfg)
    if cfg.use_boundary_refinement:   # BoundaryRefinementModule :58-149
        logits = boundary_refine(sd, p + "boundary_refiner.", logits, cfg)
    if not aux_branches:
        return logits, aux
    if cfg.use_contour_detection:     # ContourDetectionBranch :255-295
        c = p + "contour_branch.contour_branch."
        x = act_ref(norm(sd, c + "1.", conv(sd, c + "0.", shared, 1), cfg), cfg)
        x = act_ref(norm(sd, c + "4.", conv(sd, c + "3.", x, 1), cfg), cfg)
        aux["contours"] = _to_mask_size(torch.sigmoid(conv(sd, c + "6.", x, 0)), cfg)
    if cfg.use_distance_transform:    # DistanceTransformDecoder :298-344
        d = p + "distance_decoder."
        x = act_ref(norm(sd, d + "distance_head.1.", conv(sd, d + "distance_head.0.", shared, 1), cfg), cfg)
        x = residual_block(sd, d + "distance_head.3.", x, cfg, act_ref)
        dist = conv(sd, d + "distance_head.4.", x, 0)
        mask = torch.sigmoid((dist - sd[d + "threshold"]) * 10)
        aux["distance_mask"] = _to_mask_size(mask, cfg)
        aux["distance_map"] = _to_mask_size(dist, cfg)
    return logits, aux


def guided_head(sd: SD, p: str, feats: Tensor, bg_fg_mask: Tensor, cfg: PathConfig):
    """PretrainedUNetGuidedSegmentationHead.forward (rgb.py:125-218; ctor :46-123), eval mode (dropout = identity).
    Stand-alone activations come from rgb.py's factory (swish == silu), the residual blocks are the refinement ones."""
    if bg_fg_mask.shape[1] == 2:
        bg_fg_mask = bg_fg_mask[:, 1:2]                                  # :139-141 foreground channel
    fg_prob = torch.sigmoid(bg_fg_mask)
    fg_low = fg_prob
    if fg_low.shape[2:] != feats.shape[2:]:
        fg_low = F.interpolate(fg_prob, size=feats.shape[2:], mode="bilinear", align_corners=False)
    x = conv(sd, p + "input_adjust.", torch.cat([feats, fg_low], 1), 0)
    q = p + "feature_processor."
    x = act_rgb(norm(sd, q + "1.", conv(sd, q + "0.", x, 1), cfg), cfg)
    x = residual_block(sd, q + "4.", x, cfg, act_ref)
    x = residual_block(sd, q + "6.", x, cfg, act_ref)
    attention = None
    if cfg.use_attention_module:                                         # :168-173
        a = p + "attention_module."
        attention = torch.sigmoid(conv(sd, a + "2.", act_rgb(conv(sd, a + "0.", x, 0), cfg), 0))
        x = x * (attention * (0.5 + 0.5 * fg_low))
    c = p + "final_classifier."
    y = act_rgb(norm(sd, c + "1.", conv(sd, c + "0.", x, 1), cfg), cfg)
    logits = _to_mask_size(conv(sd, c + "3.", y, 0), cfg)
    mh, mw = cfg.mask_size
    if bg_fg_mask.shape[2] != mh or bg_fg_mask.shape[3] != mw:          # :188-196
        bg_fg_mask = F.interpolate(bg_fg_mask, size=(mh, mw), mode="bilinear", align_corners=False)
        fg_prob = torch.sigmoid(bg_fg_mask)
    bg_fg_logits = torch.cat([torch.log(1 - fg_prob + 1e-7), torch.log(fg_prob + 1e-7)], 1)
    aux = {"bg_fg_logits": bg_fg_logits, "target_nontarget_logits": logits[:, 1:3].clone(), "fg_prob": fg_prob,
           "pretrained_bg_fg_mask": bg_fg_mask, "attention": attention}
    return logits, aux


def uses_refined_head(cfg: PathConfig) -> bool:
    """rgb.py:683-689: any refinement flag selects RefinedHierarchicalSegmentationHead (+ feature_combiner)."""
    return bool(cfg.use_contour_detection or cfg.use_distance_transform or cfg.use_boundary_refinement or cfg.use_subpixel_conv
                or cfg.use_progressive_upsampling)


@torch.no_grad()
def forward_standard(sd: SD, images: Tensor, rois: Tensor, cfg: PathConfig):
    """HierarchicalRGBSegmentationModel.forward (rgb.py:410-439; ctor :300-408): DynamicRoIAlign(aligned=False) ->
    RGBFeatureExtractor (rgb.py:221-295) -> HierarchicalSegmentationHeadUNetV2 (..._unet.py:670-845: LayerNorm2d + ReLU
    hard-coded, EnhancedUNet 96/3) or the refined head (normalisation from the kwargs).  The reference forwards no activation
    kwarg to any sub-module of this model, so every activation is ReLU."""
    from dataclasses import replace
    rh, rw = cfg.roi_size
    sh, sw = cfg.spatial_scale
    ecfg = replace(cfg, activation_function="relu", hierarchical_base_channels=96, hierarchical_depth=3)
    x = roi_align(images, rois, rh, rw, sh, sw, False)
    roi_rgb = x
    p = "rgb_extractor.features."
    for i, base in enumerate((0, 3, 7, 11)):
        x = act_rgb(norm(sd, f"{p}{base + 1}.", conv(sd, f"{p}{base}.", x, 1), ecfg), ecfg)
        if i >= 1:
            x = residual_block(sd, f"{p}{base + 3}.", x, ecfg, act_rgb)
    if uses_refined_head(cfg):
        logits, aux = refined_head(sd, "segmentation_head.", x, ecfg)
    else:
        logits, aux = base_head(sd, "segmentation_head.", x, replace(ecfg, normalization_type="layernorm2d"))
        aux.pop("shared_features")
    aux["roi_patches"] = roi_rgb
    return logits, aux


def _rgb_feature_extractor_std(sd: SD, p: str, x: Tensor, cfg: PathConfig, rb_cfg: PathConfig) -> Tensor:
    """RGBFeatureExtractor.features (rgb.py:221-295): stages use ``cfg`` (norm + rgb-factory activation); the residual blocks
    use ``rb_cfg`` (only the batchnorm branch forwards norm/activation, every other norm gets the default LayerNorm2d+ReLU block)."""
    for i, base in enumerate((0, 3, 7, 11)):
        x = act_rgb(norm(sd, f"{p}{base + 1}.", conv(sd, f"{p}{base}.", x, 1), cfg), cfg)
        if i >= 1:
            x = residual_block(sd, f"{p}{base + 3}.", x, rb_cfg, act_rgb)
    return x


@torch.no_grad()
def forward_multiscale(sd: SD, images: Tensor, rois: Tensor, cfg: PathConfig):
    """MultiScaleRGBSegmentationModel.forward (rgb.py:866-922; ctor :780-864)."""
    from dataclasses import replace
    sh, sw = cfg.spatial_scale
    bn = cfg.normalization_type.lower() in ("batch", "batchnorm", "batchnorm2d")
    rb_cfg = cfg if bn else replace(cfg, normalization_type="layernorm2d", activation_function="relu")
    feats, first = [], None
    for name, rs in cfg.roi_sizes:
        reg = roi_align(images, rois, rs, rs, sh, sw, False)
        if first is None:
            first = reg
        f = _rgb_feature_extractor_std(sd, f"rgb_extractors.{name}.features.", reg, cfg, rb_cfg)
        if f.shape[-1] != 28:
            f = F.interpolate(f, size=(28, 28), mode="bilinear", align_corners=False)
        feats.append(f)
    if cfg.fusion_method == "concat":
        fused = torch.cat(feats, 1)
    elif cfg.fusion_method == "sum":
        fused = sum(feats)
    elif cfg.fusion_method == "adaptive":
        w = F.softmax(sd["fusion_weights"], 0)
        fused = sum(wi * fi for wi, fi in zip(w, feats))
    else:
        raise ValueError(f"Unknown fusion method: {cfg.fusion_method}")
    x = act_rgb(norm(sd, "fusion_proj.1.", conv(sd, "fusion_proj.0.", fused, 0), cfg), cfg)
    head_cfg = replace(cfg, normalization_type="layernorm2d", activation_function="relu", hierarchical_base_channels=96, hierarchical_depth=3)
    logits, aux = base_head(sd, "segmentation_head.", x, head_cfg)
    aux.pop("shared_features")
    aux["roi_patches"] = first
    return logits, aux


@torch.no_grad()
def forward(sd: SD, images: Tensor, rois: Tensor, cfg: PathConfig, full_image_logits: Tensor = None, masks_only: bool = False):
    """HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet.forward (rgb.py:729-774).  masks_only: the work of the exported
    ONNX contract (no aux-only contour / distance branches), see refined_head."""
    if cfg.multi_scale:
        return forward_multiscale(sd, images, rois, cfg)
    if not cfg.use_pretrained_unet:
        return forward_standard(sd, images, rois, cfg)
    if full_image_logits is None:
        full_image_logits = pretrained_unet_logits(sd, images, cfg)
    rh, rw = cfg.roi_size
    sh, sw = cfg.spatial_scale
    roi_masks = roi_align(full_image_logits, rois, rh, rw, sh, sw, True)
    roi_rgb = roi_align(images, rois, rh, rw, sh, sw, True)
    feats = rgb_feature_extractor(sd, roi_rgb, cfg)
    if uses_refined_head(cfg):
        comb = conv(sd, "feature_combiner.", torch.cat([feats, roi_masks], 1), 0)
        logits, aux = refined_head(sd, "segmentation_head.", comb, cfg, aux_branches=not masks_only)
    else:
        logits, aux = guided_head(sd, "segmentation_head.", feats, roi_masks, cfg)
    aux["full_image_logits"] = full_image_logits
    aux["roi_features"] = roi_masks
    aux["roi_patches"] = roi_rgb
    return logits, aux


# ----------------------------------------------------------------------------- export contract
def export_outputs(logits: Tensor, full_image_logits: Tensor):
    """RGBHierarchicalWrapper (hed/export_onnx_advanced.py:353-420): instance_masks =
    where(argmax(masks,1)==1, 1, 0); binary_masks = softmax(2ch)[:,0:1]."""
    inst = (logits.argmax(1, keepdim=True) == 1).to(logits.dtype)
    binary = F.softmax(full_image_logits, dim=1)[:, 0:1]
    return inst, binary


def mask_dilation(masks: Tensor, dilation_pixels: int = 1) -> Tensor:
    """MaskDilationModule (export_hierarchical_instance_peopleseg_onnx.py:85-141)."""
    if dilation_pixels <= 0:
        return masks
    p = F.softmax(masks, dim=1)[:, 1:2]
    d = F.max_pool2d(p, 2 * dilation_pixels + 1, 1, dilation_pixels)
    out = masks.clone()
    out[:, 1:2] = torch.where((d - p) > 0.1, masks[:, 1:2] + 2.0, masks[:, 1:2])
    return out
