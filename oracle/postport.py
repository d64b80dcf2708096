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
