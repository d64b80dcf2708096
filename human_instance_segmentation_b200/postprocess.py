"""Post-processing modules with the reference's interfaces, on the B200 stencil kernels (csrc/post.cu).

Mirrors: ``RGBHierarchicalWrapper._to_instance_masks`` (hed/export_onnx_advanced.py:360-364),
``MaskDilationModule`` (export_hierarchical_instance_peopleseg_onnx.py:85-141), ``BinaryMaskEdgeSmoothing`` /
``MultiClassEdgeSmoothing`` (hed/edge_smoothing.py:10-170), ``DirectionalEdgeSmoothing`` / ``AdaptiveEdgeSmoothing`` /
``OptimizedEdgeSmoothing`` (export_edge_smoothing_onnx.py:63-318), ``BilateralFilter`` / ``FastBilateralFilter`` /
``EdgePreservingFilter`` / ``BinaryMaskBilateralFilter`` / ``MorphologicalBilateralFilter`` (hed/bilateral_filter.py:9-501),
the NEAREST paste-back of test_hierarchical_instance_peopleseg_onnx.py, and ``MaskCleanup`` -- edge smoothing + binary
bilateral filter fused into one shared-memory pass (BASELINE config 5).
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


@_lib.on_tensor_device
def instance_masks(logits: torch.Tensor, score_threshold: float = 0.0, as_uint8: bool = False, dilation_pixels: int = 0) -> torch.Tensor:
    """[N,3,H,W] logits -> [N,1,H,W] fp32 in {0,1} (or [N,H,W] uint8).  ``dilation_pixels > 0`` applies MaskDilationModule to
    the logits first, inside the same kernel (== ``instance_masks(MaskDilationModule(d)(logits))`` bit for bit, without the
    dilated tensor going through HBM)."""
    x = _cuda_f32(logits, "instance_masks")
    n, c, h, w = x.shape
    if c != 3:
        raise ValueError("instance_masks expects 3-class logits [N,3,H,W]")
    L = _lib.load()
    out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device) if as_uint8 else torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
    ptrs = (None, out.data_ptr()) if as_uint8 else (out.data_ptr(), None)
    if dilation_pixels > 0:
        _lib.check(L.his_post_dilate_instance_mask(x.data_ptr(), n, h, w, int(dilation_pixels), float(score_threshold), *ptrs, _stream(x)))
    else:
        _lib.check(L.his_post_instance_mask(x.data_ptr(), n, h, w, float(score_threshold), *ptrs, _stream(x)))
    return out


class MaskDilationModule(nn.Module):
    def __init__(self, dilation_pixels: int = 1):
        super().__init__()
        self.dilation_pixels = dilation_pixels

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, masks: torch.Tensor) -> torch.Tensor:
        if self.dilation_pixels <= 0:          # export_hierarchical_instance_peopleseg_onnx.py:97-98 returns the input
            return masks
        x = _cuda_f32(masks, "MaskDilationModule")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"MaskDilationModule expects [N,3,H,W] logits (bg, target, non-target), got {tuple(x.shape)}")
        n, c, h, w = x.shape
        if n == 0:
            return x.clone()
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

    @_lib.on_tensor_device
    def forward(self, mask: torch.Tensor) -> torch.Tensor:
        x4, shape = _as_planes(mask)
        x = _cuda_f32(x4, "BinaryMaskEdgeSmoothing")
        b, c, h, w = x.shape
        out = torch.empty_like(x)
        _run_planes(_lib.load().his_post_edge_smooth_tiled, x, out, float(self.threshold), float(self.blur_strength))
        return out.to(mask.dtype).reshape(shape)


def _run_planes(fn, x: torch.Tensor, out: torch.Tensor, *args, extra_in=(), extra_out=()):
    """Calls a tiled stencil entry point over the [B*C] planes of ``x`` (one launch whatever the plane count: the plane index
    rides on gridDim.x)."""
    b, c, h, w = x.shape
    if b * c == 0:
        return
    _lib.check(fn(x.data_ptr(), *[t.data_ptr() for t in extra_in], b * c, h, w, *args, *[t.data_ptr() for t in extra_out], out.data_ptr(), _stream(x)))


class DirectionalEdgeSmoothing(nn.Module):
    """export_edge_smoothing_onnx.py:63-154."""

    def __init__(self, num_directions: int = 4):
        super().__init__()
        self.num_directions = num_directions

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, mask: torch.Tensor) -> torch.Tensor:
        x = _cuda_f32(mask, "DirectionalEdgeSmoothing")
        if x.dim() != 4 or x.shape[1] != 1:
            raise RuntimeError("DirectionalEdgeSmoothing expects [B,1,H,W] (the reference's conv kernels have one input channel)")
        out = torch.empty_like(x)
        _run_planes(_lib.load().his_post_edge_directional, x, out)
        return out


class AdaptiveEdgeSmoothing(nn.Module):
    """export_edge_smoothing_onnx.py:157-218; the three parameters are per-image tensors [B,1]."""

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, mask, blur_strength, edge_sensitivity, final_threshold) -> torch.Tensor:
        x = _cuda_f32(mask, "AdaptiveEdgeSmoothing")
        if x.dim() != 4 or x.shape[1] != 1:
            raise RuntimeError("AdaptiveEdgeSmoothing expects [B,1,H,W]")
        b, _, h, w = x.shape
        prm = [t.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous() for t in (blur_strength, edge_sensitivity, final_threshold)]
        if any(t.numel() != b for t in prm):
            raise RuntimeError("AdaptiveEdgeSmoothing: parameters must hold one value per image ([B,1])")
        out = torch.empty_like(x)
        _lib.check(_lib.load().his_post_edge_adaptive(x.data_ptr(), b, h, w, prm[0].data_ptr(), prm[1].data_ptr(), prm[2].data_ptr(),
                                                      out.data_ptr(), _stream(x)))
        return out


class OptimizedEdgeSmoothing(nn.Module):
    """export_edge_smoothing_onnx.py:221-318.  ``use_fp16`` reproduces the exported FP16 graph (every operator output rounded to
    half) and returns a half tensor, like the reference."""

    def __init__(self, use_fp16: bool = True):
        super().__init__()
        self.use_fp16 = use_fp16

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, mask: torch.Tensor) -> torch.Tensor:
        x = _cuda_f32(mask, "OptimizedEdgeSmoothing")
        if x.dim() != 4 or x.shape[1] != 1:
            raise RuntimeError("OptimizedEdgeSmoothing expects [B,1,H,W]")
        out = torch.empty_like(x)
        _run_planes(_lib.load().his_post_edge_optimized, x, out, 1 if self.use_fp16 else 0)
        return out.half() if self.use_fp16 else out


class MultiClassEdgeSmoothing:
    """hed/edge_smoothing.py:93-170 (``device`` is accepted for signature parity; tensors must already be CUDA tensors)."""

    def __init__(self, threshold: float = 0.5, blur_strength: float = 3.0, iterations: int = 1, device: str = "cuda"):
        self.smoother = BinaryMaskEdgeSmoothing(threshold, blur_strength)
        self.iterations = iterations
        self.device = device

    @torch.no_grad()

    @_lib.on_tensor_device
    def smooth_predictions(self, predictions, apply_softmax: bool = False):
        """``predictions``: CUDA tensor, or (like the reference, edge_smoothing.py:120-124) a numpy array, which is uploaded to
        ``self.device`` and returned as a numpy array."""
        import numpy as np
        if isinstance(predictions, np.ndarray):
            out = self.smooth_predictions(torch.from_numpy(predictions).to(self.device), apply_softmax)
            return out.cpu().numpy()
        x = _cuda_f32(predictions, "MultiClassEdgeSmoothing")
        squeeze = x.dim() == 3
        if squeeze:
            x = x[None]
        b, c, h, w = x.shape
        L = _lib.load()
        masks = torch.empty_like(x)
        # argmax is invariant under softmax, so the softmax only matters for the thresholded (C != 3) flavour
        _lib.check(L.his_post_class_masks(x.data_ptr(), b, c, h, w, 1 if c == 3 else 0, 1 if apply_softmax else 0, masks.data_ptr(), _stream(x)))
        for _ in range(self.iterations):
            masks = self.smoother(masks)
        return masks[0] if (squeeze or b == 1) else masks


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


class BilateralFilter(nn.Module):
    """hed/bilateral_filter.py:9-113 (the reference walks every pixel in Python; here one tiled kernel)."""

    def __init__(self, kernel_size: int = 5, sigma_spatial: float = 1.0, sigma_range: float = 0.1):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("Kernel size must be odd")
        self.kernel_size, self.sigma_spatial, self.sigma_range, self.padding = kernel_size, sigma_spatial, sigma_range, kernel_size // 2
        coords = torch.arange(kernel_size, dtype=torch.float32) - (kernel_size - 1) / 2
        y, x = coords.view(-1, 1).expand(kernel_size, kernel_size), coords.view(1, -1).expand(kernel_size, kernel_size)
        self.register_buffer("spatial_kernel", torch.exp(-(x ** 2 + y ** 2) / (2 * sigma_spatial ** 2)))

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _cuda_f32(x, "BilateralFilter")
        out = torch.empty_like(x)
        sk = self.spatial_kernel.to(x.device).contiguous()
        _run_planes(_lib.load().his_post_bilateral_exact, x, out, sk.data_ptr(), self.kernel_size, float(self.sigma_range))
        return out


class FastBilateralFilter(nn.Module):
    """hed/bilateral_filter.py:116-216."""

    def __init__(self, kernel_size: int = 5, sigma_spatial: float = 1.0, sigma_range: float = 0.1, num_iterations: int = 2):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("Kernel size must be odd")
        self.kernel_size, self.sigma_spatial, self.sigma_range, self.num_iterations = kernel_size, sigma_spatial, sigma_range, num_iterations
        self.padding = kernel_size // 2
        coords = torch.arange(kernel_size, dtype=torch.float32) - (kernel_size - 1) / 2
        k1 = torch.exp(-coords ** 2 / (2 * sigma_spatial ** 2))
        k1 = k1 / k1.sum()
        self.register_buffer("kernel_h", k1.view(1, 1, 1, kernel_size))
        self.register_buffer("kernel_v", k1.view(1, 1, kernel_size, 1))

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _cuda_f32(x, "FastBilateralFilter")
        if self.num_iterations < 1:
            return x.clone()
        out, ws = torch.empty_like(x), torch.empty_like(x)
        k1 = self.kernel_h.to(x.device).reshape(-1).contiguous()
        _run_planes(_lib.load().his_post_bilateral_fast, x, out, k1.data_ptr(), self.kernel_size, float(self.sigma_range), int(self.num_iterations),
                    extra_out=(ws,))
        return out


class EdgePreservingFilter(nn.Module):
    """Guided filter, hed/bilateral_filter.py:219-296."""

    def __init__(self, radius: int = 2, eps: float = 0.01):
        super().__init__()
        self.radius, self.eps, self.kernel_size = radius, eps, 2 * radius + 1

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, x: torch.Tensor, guide: torch.Tensor = None) -> torch.Tensor:
        x = _cuda_f32(x, "EdgePreservingFilter")
        g = x if guide is None else _cuda_f32(guide, "EdgePreservingFilter")
        if g.shape != x.shape:
            raise RuntimeError("EdgePreservingFilter: guide must have the input's shape")
        out, a, b = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        _run_planes(_lib.load().his_post_guided_filter, x, out, int(self.radius), float(self.eps), extra_in=(g,), extra_out=(a, b))
        return out


class MaskCleanup(nn.Module):
    """``BinaryMaskBilateralFilter(BinaryMaskEdgeSmoothing(mask))`` in one kernel: the post-processing chain of BASELINE
    config 5 on full-image masks, read once and written once (bit-identical to running the two modules back to back)."""

    def __init__(self, es_threshold: float = 0.5, blur_strength: float = 3.0, kernel_size: int = 7, sigma_spatial: float = 1.5,
                 threshold: float = 0.5, num_iterations: int = 2):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("Kernel size must be odd")
        self.es_threshold, self.blur_strength = es_threshold, blur_strength
        self.kernel_size, self.threshold, self.num_iterations = kernel_size, threshold, num_iterations
        self.register_buffer("gaussian_kernel", _gauss2d(kernel_size, sigma_spatial, False).view(1, 1, kernel_size, kernel_size))

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """fp32 masks [B,C,H,W] -> fp32 {0,1}; uint8 masks (0 / 1) -> uint8 through the byte variant of the same kernel (the result
        is that of the fp32 call on ``x.float()``, with a quarter of the bytes moved)."""
        g = self.gaussian_kernel.to(x.device).contiguous()
        if x.dtype == torch.uint8:
            if not x.is_cuda:
                raise _lib.HisError("MaskCleanup: CUDA tensor required (no CPU fallback)")
            x = x.contiguous()
            out = torch.empty_like(x) if out is None else out
            fn = _lib.load().his_post_mask_cleanup_fused_u8
        else:
            x = _cuda_f32(x, "MaskCleanup")
            out = torch.empty_like(x) if out is None else out
            fn = _lib.load().his_post_mask_cleanup_fused
        if out.dtype != x.dtype or out.shape != x.shape or not out.is_contiguous():
            raise ValueError("MaskCleanup: out must be a contiguous tensor of the input's shape and dtype")
        _run_planes(fn, x, out, float(self.es_threshold), float(self.blur_strength), g.data_ptr(), self.kernel_size, int(self.num_iterations),
                    float(self.threshold))
        return out


class BinaryMaskBilateralFilter(nn.Module):
    def __init__(self, kernel_size: int = 7, sigma_spatial: float = 1.5, threshold: float = 0.5, num_iterations: int = 2):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("Kernel size must be odd")
        self.kernel_size, self.sigma_spatial, self.threshold, self.num_iterations = kernel_size, sigma_spatial, threshold, num_iterations
        self.register_buffer("gaussian_kernel", _gauss2d(kernel_size, sigma_spatial, False).view(1, 1, kernel_size, kernel_size))

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _cuda_f32(x, "BinaryMaskBilateralFilter")
        b, c, h, w = x.shape
        g = self.gaussian_kernel.to(x.device).contiguous()
        out = torch.empty_like(x)
        if self.kernel_size <= 9 and 1 <= self.num_iterations and self.num_iterations * (self.kernel_size // 2) <= 8:
            # all iterations in one shared-memory pass
            _run_planes(_lib.load().his_post_binary_bilateral_tiled, x, out, g.data_ptr(), self.kernel_size, int(self.num_iterations),
                        float(self.threshold))
            return out
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

    @_lib.on_tensor_device
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


@_lib.on_tensor_device
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
    if n:
        _lib.check(_lib.load().his_post_paste(m.data_ptr(), n, mh, mw, r.data_ptr(), canvas.data_ptr(), batch_size, height, width, _stream(m)))
    return canvas
