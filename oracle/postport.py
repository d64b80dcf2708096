"""Oracle (test infrastructure): CPU restatement of the post-processing family.

torch fp32 on CPU, the same operators the reference modules call, so thresholds see the reference's arithmetic.
Pinned against the reference modules/functions themselves by tests/golden/post.npz (oracle/make_golden_post.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def instance_mask(logits: torch.Tensor, score_threshold: float = 0.0) -> torch.Tensor:
    """hed/export_onnx_advanced.py:360-364 (argmax==1); with a threshold: test_hierarchical_instance_peopleseg_onnx.py:250-262
    ((argmax(softmax)==1) & (max prob > thr))."""
    on = logits.argmax(1, keepdim=True) == 1
    if score_threshold > 0:
        on &= F.softmax(logits, 1).max(1, keepdim=True)[0] > score_threshold
    return on.float()


def edge_smooth(mask: torch.Tensor, threshold: float = 0.5, blur_strength: float = 3.0) -> torch.Tensor:
    """BinaryMaskEdgeSmoothing.forward, hed/edge_smoothing.py:35-90, for [B,C,H,W]."""
    lap = torch.tensor([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]], dtype=torch.float32).view(1, 1, 3, 3)
    g = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32).view(1, 1, 3, 3) / 16
    outs = []
    for c in range(mask.shape[1]):
        m = mask[:, c:c + 1].float()
        w = torch.sigmoid(F.conv2d(m, lap, padding=1).abs() * blur_strength)
        sm = m * (1 - w) + F.conv2d(m, g, padding=1) * w
        outs.append((sm > threshold).to(mask.dtype))
    return torch.cat(outs, 1)


def _gauss2d(k, sigma, outer):
    coords = torch.arange(k, dtype=torch.float32) - (k - 1) / 2
    if outer:
        k1 = torch.exp(-coords ** 2 / (2 * sigma ** 2)); k1 = k1 / k1.sum()
        return (k1.view(-1, 1) * k1.view(1, -1)).view(1, 1, k, k)
    y = coords.view(-1, 1).expand(k, k); x = coords.view(1, -1).expand(k, k)
    g = torch.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).view(1, 1, k, k)


def binary_bilateral(x: torch.Tensor, kernel_size=7, sigma_spatial=1.5, threshold=0.5, num_iterations=2, return_soft=False):
    """BinaryMaskBilateralFilter.forward, hed/bilateral_filter.py:349-406."""
    g = _gauss2d(kernel_size, sigma_spatial, False)
    pad = kernel_size // 2
    x = torch.clamp(x, 0, 1)
    chans = []
    for c in range(x.shape[1]):
        m = x[:, c:c + 1]
        for _ in range(num_iterations):
            f = F.conv2d(m, g, padding=pad)
            var = torch.clamp(F.conv2d(m ** 2, g, padding=pad) - f ** 2, min=0)
            ew = torch.exp(-var * 10)
            m = ew * f + (1 - ew) * m
        chans.append(m)
    soft = torch.cat(chans, 1)
    return soft if return_soft else (soft > threshold).float()


def morph_bilateral(x: torch.Tensor, kernel_size=5, sigma=1.0, morph_size=3, return_soft=False):
    """MorphologicalBilateralFilter.forward, hed/bilateral_filter.py:478-501."""
    k2 = _gauss2d(kernel_size, sigma, True)
    mp = morph_size // 2
    x = torch.clamp(x, 0, 1)
    er = -F.max_pool2d(-x, morph_size, 1, mp)
    op = F.max_pool2d(er, morph_size, 1, mp)
    f = torch.cat([F.conv2d(op[:, c:c + 1], k2, padding=kernel_size // 2) for c in range(x.shape[1])], 1) if x.shape[1] != 1 \
        else F.conv2d(op, k2, padding=kernel_size // 2)
    di = F.max_pool2d(f, morph_size, 1, mp)
    cl = -F.max_pool2d(-di, morph_size, 1, mp)
    return cl if return_soft else (cl > 0.5).float()


# ----------------------------------------------------------------------------- a16 / a17 variants
def _conv(x, k, pad):
    return F.conv2d(x, k, padding=pad)


def directional_edge_smooth(mask: torch.Tensor) -> torch.Tensor:
    """DirectionalEdgeSmoothing.forward, export_edge_smoothing_onnx.py:63-154 ([B,1,H,W] float masks)."""
    t = lambda v: torch.tensor(v, dtype=torch.float32).view(1, 1, len(v), len(v[0]))
    sx, sy = t([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]), t([[-1, -2, -1], [0, 0, 0], [1, 2, 1]])
    hb, vb = t([[0.1, 0.2, 0.4, 0.2, 0.1]]), t([[0.1], [0.2], [0.4], [0.2], [0.1]])
    d1, d2 = t([[0.1, 0, 0], [0, 0.8, 0], [0, 0, 0.1]]), t([[0, 0, 0.1], [0, 0.8, 0], [0.1, 0, 0]])
    ex, ey = _conv(mask, sx, 1), _conv(mask, sy, 1)
    mag = torch.sqrt(ex ** 2 + ey ** 2 + 1e-8)
    ang = torch.atan2(ey, ex)
    bh, bv = _conv(mask, hb, (0, 2)), _conv(mask, vb, (2, 0))
    b1, b2 = _conv(mask, d1, 1), _conv(mask, d2, 1)
    wh, wv = torch.cos(ang) ** 2, torch.sin(ang) ** 2
    w1, w2 = torch.cos(ang - np.pi / 4) ** 2 * 0.5, torch.cos(ang + np.pi / 4) ** 2 * 0.5
    ws = wh + wv + w1 + w2 + 1e-8
    blurred = bh * (wh / ws) + bv * (wv / ws) + b1 * (w1 / ws) + b2 * (w2 / ws)
    em = torch.sigmoid(mag * 3.0)
    return ((mask * (1 - em) + blurred * em) > 0.5).float()


def adaptive_edge_smooth(mask, blur_strength, edge_sensitivity, final_threshold) -> torch.Tensor:
    """AdaptiveEdgeSmoothing.forward, export_edge_smoothing_onnx.py:157-218 (per-image [B,1] parameters)."""
    B = mask.shape[0]
    lap = torch.tensor([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]], dtype=torch.float32).view(1, 1, 3, 3)
    edges = _conv(mask, lap, 1).abs()
    em = (edges > 0.5 * edge_sensitivity.view(B, 1, 1, 1)).float()
    base = _conv(mask, torch.ones(1, 1, 5, 5) / 25, 2)
    bf = blur_strength.view(B, 1, 1, 1) / 3.0
    sm = mask * (1 - bf) + base * bf
    return ((mask * (1 - em) + sm * em) > final_threshold.view(B, 1, 1, 1)).float()


def optimized_edge_smooth(mask: torch.Tensor) -> torch.Tensor:
    """OptimizedEdgeSmoothing.forward with use_fp16=False, export_edge_smoothing_onnx.py:221-318 (separable 5-tap
    binomial blur, clamp((3|lap|+0.5)/2, 0, 1) in place of the sigmoid)."""
    lap = torch.tensor([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]], dtype=torch.float32).view(1, 1, 3, 3)
    g = torch.tensor([0.0625, 0.25, 0.375, 0.25, 0.0625], dtype=torch.float32)
    e = _conv(mask, lap, 1).abs() * 3.0
    bl = _conv(_conv(mask, g.view(1, 1, 1, 5), (0, 2)), g.view(1, 1, 5, 1), (2, 0))
    em = torch.clamp((e + 0.5) * 0.5, 0, 1)
    return ((mask * (1 - em) + bl * em) > 0.5).float()


def multiclass_edge_smooth(pred: torch.Tensor, threshold=0.5, blur_strength=3.0, iterations=1, apply_softmax=False):
    """MultiClassEdgeSmoothing.smooth_predictions, hed/edge_smoothing.py:93-170, for [B,C,H,W]."""
    if apply_softmax:
        pred = torch.softmax(pred, 1)
    C = pred.shape[1]
    outs = []
    for c in range(C):
        m = (pred.argmax(1) == c).float() if C == 3 else (pred[:, c] > 0.5).float()
        for _ in range(iterations):
            m = edge_smooth(m[:, None], threshold, blur_strength)[:, 0]
        outs.append(m)
    return torch.stack(outs, 1)


def exact_bilateral(x: torch.Tensor, kernel_size=5, sigma_spatial=1.0, sigma_range=0.1) -> torch.Tensor:
    """BilateralFilter.forward, hed/bilateral_filter.py:9-113, vectorised with unfold (same arithmetic per pixel)."""
    k, pad = kernel_size, kernel_size // 2
    coords = torch.arange(k, dtype=torch.float32) - (k - 1) / 2
    sp = torch.exp(-(coords.view(1, -1) ** 2 + coords.view(-1, 1) ** 2) / (2 * sigma_spatial ** 2)).reshape(1, 1, k * k, 1)
    B, C, H, W = x.shape
    xp = F.pad(x, [pad] * 4, mode="reflect")
    patches = F.unfold(xp.reshape(B * C, 1, H + 2 * pad, W + 2 * pad), k).reshape(B, C, k * k, H * W)
    center = x.reshape(B, C, 1, H * W)
    w = sp * torch.exp(-(patches - center) ** 2 / (2 * sigma_range ** 2))
    wn = w / (w.sum(2, keepdim=True) + 1e-8)
    return (patches * wn).sum(2).reshape(B, C, H, W)


def fast_bilateral(x: torch.Tensor, kernel_size=5, sigma_spatial=1.0, sigma_range=0.1, num_iterations=2) -> torch.Tensor:
    """FastBilateralFilter.forward, hed/bilateral_filter.py:116-216."""
    pad = kernel_size // 2
    coords = torch.arange(kernel_size, dtype=torch.float32) - (kernel_size - 1) / 2
    k1 = torch.exp(-coords ** 2 / (2 * sigma_spatial ** 2)); k1 = k1 / k1.sum()
    kh, kv = k1.view(1, 1, 1, -1), k1.view(1, 1, -1, 1)
    chans = []
    for c in range(x.shape[1]):
        ch = x[:, c:c + 1]
        for _ in range(num_iterations):
            f = _conv(_conv(ch, kh, (0, pad)), kv, (pad, 0))
            sq = _conv(_conv(ch ** 2, kh, (0, pad)), kv, (pad, 0))
            ew = torch.exp(-torch.clamp(sq - f ** 2, min=0) / (2 * sigma_range ** 2))
            ch = ew * f + (1 - ew) * ch
        chans.append(ch)
    return torch.cat(chans, 1)


def edge_preserving(x: torch.Tensor, guide: torch.Tensor = None, radius=2, eps=0.01) -> torch.Tensor:
    """EdgePreservingFilter.forward (guided filter), hed/bilateral_filter.py:219-296; box filters are zero padded."""
    k = 2 * radius + 1
    box = lambda t: torch.cat([_conv(t[:, c:c + 1], torch.ones(1, 1, k, k) / (k * k), radius) for c in range(t.shape[1])], 1)
    g = x if guide is None else guide
    mx, mg = box(x), box(g)
    a = (box(x * g) - mx * mg) / ((box(g * g) - mg * mg) + eps)
    b = mx - a * mg
    return box(a) * g + box(b)


def nearest_index(dst: int, src: int) -> np.ndarray:
    """cv2.resize(INTER_NEAREST) source index: min(floor(x * (1/(dst/src))), src-1) in double."""
    scale = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * scale).astype(np.int64), src - 1)


def paste_back(masks_u8: np.ndarray, rois: np.ndarray, batch: int, height: int, width: int) -> np.ndarray:
    """test_hierarchical_instance_peopleseg_onnx.py:144-161 (int() truncation of the fp32 product), :264-278 (NEAREST
    resize), :369-374 (full[y1:y2, x1:x2] = mask; later instances overwrite) -> int32 label canvas [B,H,W]."""
    canvas = np.zeros((batch, height, width), np.int32)
    rois = rois.astype(np.float32)
    for i in range(masks_u8.shape[0]):
        b = int(rois[i, 0])
        x1, y1 = int(rois[i, 1] * np.float32(width)), int(rois[i, 2] * np.float32(height))
        x2, y2 = int(rois[i, 3] * np.float32(width)), int(rois[i, 4] * np.float32(height))
        rw, rh = x2 - x1, y2 - y1
        if rw <= 0 or rh <= 0 or not (0 <= b < batch):
            continue
        m = masks_u8[i][nearest_index(rh, masks_u8.shape[1])][:, nearest_index(rw, masks_u8.shape[2])]
        ys, xs = np.nonzero(m)
        ys, xs = ys + y1, xs + x1
        ok = (ys >= 0) & (ys < height) & (xs >= 0) & (xs < width)
        canvas[b, ys[ok], xs[ok]] = i + 1
    return canvas
