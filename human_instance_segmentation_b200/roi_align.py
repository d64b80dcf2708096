"""``DynamicRoIAlign`` with the reference's interface (hed/dynamic_roi_align.py:10-171), on the B200 gather kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as _lib


class DynamicRoIAlign(nn.Module):
    """Same constructor, attributes (``spatial_scale``, ``spatial_scale_h/_w``, ``aligned`` -- the exporter mutates
    them, hed/export_onnx_advanced.py:80-98) and ``forward(input_feature_map, rois, output_height, output_width)``
    -> ``[K, C, oh, ow]`` fp32.  ``sampling_ratio`` is accepted and ignored, as in the reference (:23-24)."""

    def __init__(self, spatial_scale=(640, 640), sampling_ratio=-1, aligned=False):
        super().__init__()
        if isinstance(spatial_scale, (list, tuple)):
            assert len(spatial_scale) == 2, "spatial_scale tuple must have 2 elements (height, width)"
            self.spatial_scale = spatial_scale
            self.spatial_scale_h, self.spatial_scale_w = spatial_scale[0], spatial_scale[1]
        else:
            self.spatial_scale = spatial_scale
            self.spatial_scale_h = self.spatial_scale_w = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    @torch.no_grad()

    @_lib.on_tensor_device
    def forward(self, input_feature_map: torch.Tensor, rois: torch.Tensor, output_height, output_width) -> torch.Tensor:
        if isinstance(output_width, (list, tuple)):
            output_width = output_width[0]
        if isinstance(output_height, (list, tuple)):
            output_height = output_height[0]
        oh, ow = int(output_height), int(output_width)
        if not input_feature_map.is_cuda:
            raise _lib.HisError("DynamicRoIAlign (B200) needs CUDA tensors; there is no CPU fallback")
        L = _lib.load()
        feat = input_feature_map
        if feat.dtype not in (torch.float32, torch.float16):
            feat = feat.float()
        rois = rois.to(device=feat.device, dtype=torch.float32).contiguous()
        B, C, H, W = feat.shape
        K = rois.shape[0]
        out = torch.empty((K, C, oh, ow), dtype=torch.float32, device=feat.device)
        sN, sC, sH, sW = feat.stride()
        stream = torch.cuda.current_stream(feat.device).cuda_stream
        _lib.check(L.his_roi_align(feat.data_ptr(), 1 if feat.dtype == torch.float16 else 0, sN, sC, sH, sW, B, C, H, W,
                                   rois.data_ptr(), K, oh, ow, float(self.spatial_scale_h), float(self.spatial_scale_w),
                                   1 if self.aligned else 0, None, 0, out.data_ptr(), 0, stream), "his_roi_align")
        return out
