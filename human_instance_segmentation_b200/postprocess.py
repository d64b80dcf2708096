"""Post-processing modules with the reference's interfaces, on the B200 stencil kernels (csrc/post.cu).

Mirrors: ``RGBHierarchicalWrapper._to_instance_masks`` (hed/export_onnx_advanced.py:360-364),
``MaskDilationModule`` (export_hierarchical_instance_peopleseg_onnx.py:85-141), ``BinaryMaskEdgeSmoothing``
(hed/edge_smoothing.py:10-90), ``BinaryMaskBilateralFilter`` / ``MorphologicalBilateralFilter``
(hed/bilateral_filter.py:299-501) and the NEAREST paste-back of test_hierarchical_instance_peopleseg_onnx.py.
All take/return CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as _lib


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.HisError(f"{what}: CUDA tensor required (no CPU fallback)")
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


@torch.no_grad()
def instance_masks(logits: torch.Tensor, score_threshold: float = 0.0, as_uint8: bool = False) -> torch.Tensor:
    """[N,3,H,W] logits -> [N,1,H,W] fp32 in {0,1} (or [N,H,W] uint8)."""
    x = _cuda_f32(logits, "instance_masks")
    n, c, h, w = x.shape
    if c != 3:
        raise ValueError("instance_masks expects 3-class logits [N,3,H,W]")
    L = _lib.load()
    if as_uint8:
        out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
        _lib.check(L.his_post_instance_mask(x.data_ptr(), n, h, w, float(score_threshold), None, out.data_ptr(), _stream(x)))
    else:
        out = torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
        _lib.check(L.his_post_instance_mask(x.data_ptr(), n, h, w, float(score_threshold), out.data_ptr(), None, _stream(x)))
    return out


class MaskDilationModule(nn.Module):
    def __init__(self, dilation_pixels: int = 1):
        super().__init__()
        self.dilation_pixels = dilation_pixels

    @torch.no_grad()
    def forward(self, masks: torch.Tensor) -> torch.Tensor:
        if self.dilation_pixels == 0:
            return masks
        x = _cuda_f32(masks, "MaskDilationModule")
        n, c, h, w = x.shape
        out = torch.empty_like(x)
        _lib.check(_lib.load().his_post_dilate_logits(x.data_ptr(), n, h, w, int(self.dilation_pixels), out.data_ptr(), _stream(x)))
        return out


class ModelWithDilation(nn.Module):
    """export_hierarchical_instance_peopleseg_onnx.py:144-181."""

    def __init__(self, base_model: nn.Module, dilation_pixels: int = 1):
        super().__init__()
        self.base_model = base_model
        self.dilation = MaskDilationModule(dilation_pixels) if dilation_pixels > 0 else None

    def forward(self, images, rois):
        output = self.base_model(images, rois)
        if isinstance(output, tuple):
            masks = output[0]
            if self.dilation is not None:
                masks = self.dilation(masks)
            return (masks,) + output[1:] if len(output) > 1 else masks
        return self.dilation(output) if self.dilation is not None else output


def _as_planes(mask: torch.Tensor):
    shape = mask.shape
    x = mask
    if x.dim() == 2:
        x = x[None, None]
    elif x.dim() == 3:
        x = x[None]
    return x, shape


class BinaryMaskEdgeSmoothing(nn.Module):
    def __init__(self, threshold: float = 0.5, blur_strength: float = 3.0):
        super().__init__()
        self.threshold, self.blur_strength = threshold, blur_strength

    @torch.no_grad()
    def forward(self, mask: torch.Tensor) -> torch.Tensor:
        x4, shape = _as_planes(mask)
        x = _cuda_f32(x4, "BinaryMaskEdgeSmoothing")
        b, c, h, w = x.shape
        out = torch.empty_like(x)
        _lib.check(_lib.load().his_post_edge_smooth(x.data_ptr(), b * c, h, w, float(self.threshold), float(self.blur_strength),
                                                    out.data_ptr(), _stream(x)))
        return out.to(mask.dtype).reshape(shape)


def _gauss2d(k: int, sigma: float, outer: bool) -> torch.Tensor:
    """The reference builds its kernels with torch fp32 arithmetic at construction time (bilateral_filter.py:333-346,
    :438-447); constants are reproduced the same way (host side, once)."""
    coords = torch.arange(k, dtype=torch.float32) - (k - 1) / 2
    if outer:
        k1 = torch.exp(-coords ** 2 / (2 * sigma ** 2))
        k1 = k1 / k1.sum()
        return (k1.view(-1, 1) * k1.view(1, -1)).contiguous()
    y = coords.view(-1, 1).expand(k, k)
    x = coords.view(1, -1).expand(k, k)
    g = torch.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).contiguous()


class BinaryMaskBilateralFilter(nn.Module):
    def __init__(self, kernel_size: int = 7, sigma_spatial: float = 1.5, threshold: float = 0.5, num_iterations: int = 2):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("Kernel size must be odd")
        self.kernel_size, self.sigma_spatial, self.threshold, self.num_iterations = kernel_size, sigma_spatial, threshold, num_iterations
        self.register_buffer("gaussian_kernel", _gauss2d(kernel_size, sigma_spatial, False).view(1, 1, kernel_size, kernel_size))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _cuda_f32(x, "BinaryMaskBilateralFilter")
        b, c, h, w = x.shape
        g = self.gaussian_kernel.to(x.device).contiguous()
        out = torch.empty_like(x)
        ws0, ws1 = torch.empty_like(x), torch.empty_like(x)
        _lib.check(_lib.load().his_post_binary_bilateral(x.data_ptr(), b * c, h, w, g.data_ptr(), self.kernel_size, int(self.num_iterations),
                                                         float(self.threshold), ws0.data_ptr(), ws1.data_ptr(), out.data_ptr(), _stream(x)))
        return out


class MorphologicalBilateralFilter(nn.Module):
    def __init__(self, kernel_size: int = 5, sigma: float = 1.0, morph_size: int = 3):
        super().__init__()
        self.kernel_size, self.sigma, self.morph_size = kernel_size, sigma, morph_size
        self.register_buffer("bilateral_kernel", _gauss2d(kernel_size, sigma, True).view(1, 1, kernel_size, kernel_size))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _cuda_f32(x, "MorphologicalBilateralFilter")
        b, c, h, w = x.shape
        g = self.bilateral_kernel.to(x.device).contiguous()
        out = torch.empty_like(x)
        ws0, ws1 = torch.empty_like(x), torch.empty_like(x)
        _lib.check(_lib.load().his_post_morph_bilateral(x.data_ptr(), b * c, h, w, g.data_ptr(), self.kernel_size, self.morph_size,
                                                        ws0.data_ptr(), ws1.data_ptr(), out.data_ptr(), _stream(x)))
        return out


@torch.no_grad()
def paste_masks(masks_u8: torch.Tensor, rois: torch.Tensor, batch_size: int, height: int, width: int) -> torch.Tensor:
    """ROI masks [N,mh,mw] uint8 + rois [N,5] -> int32 label canvas [B,H,W]: 0 = background, i+1 = last ROI i covering
    the pixel (the reference pastes instances in order, later ones overwrite)."""
    if not masks_u8.is_cuda:
        raise _lib.HisError("paste_masks: CUDA tensor required (no CPU fallback)")
    m = masks_u8.contiguous()
    if m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    r = rois.to(device=m.device, dtype=torch.float32).contiguous()
    n, mh, mw = m.shape
    canvas = torch.zeros((batch_size, height, width), dtype=torch.int32, device=m.device)
    L = _lib.load()
    for s in range(0, n, 65535):
        e = min(n, s + 65535)
        if s == 0:
            _lib.check(L.his_post_paste(m.data_ptr(), e, mh, mw, r.data_ptr(), canvas.data_ptr(), batch_size, height, width, _stream(m)))
        else:
            raise _lib.HisError("paste_masks: more than 65535 ROIs per call; chunk the batch")
    return canvas
