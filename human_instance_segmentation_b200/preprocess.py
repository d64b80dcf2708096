"""Input path of the demo / validation scripts on the GPU:
``prepare_image`` (test_hierarchical_instance_peopleseg_onnx.py:170-196) after the file decode -- BGR->RGB, ``cv2.resize`` (uint8,
INTER_LINEAR), ``/255``, HWC->CHW -- as one kernel, bit-exact with OpenCV's 8-bit bilinear arithmetic, and ``normalize_bbox``
(:118-141) for the ROI rows.  JPEG decoding stays on the host (no nvJPEG in this build)."""
from __future__ import annotations

from functools import lru_cache
from typing import Sequence, Tuple

import numpy as np
import torch

from . import lib as _lib


@lru_cache(maxsize=64)
def _tables(dn: int, sn: int, vertical: bool) -> np.ndarray:
    """{index0, index1, weight0, weight1} of cv2's INTER_LINEAR for a dn-long axis resized from sn: fx in float from a double
    scale 1/(dn/sn); 11-bit weights by round-half-even; horizontally the weight is forced to (1, 0) where a tap leaves the image,
    vertically only the two row indices are clipped."""
    scale = 1.0 / (float(dn) / float(sn))
    d = np.arange(dn)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    fx = (fx - sx).astype(np.float32)
    if not vertical:
        lo = sx < 0
        fx[lo] = 0
        sx[lo] = 0
        hi = sx >= sn - 1
        fx[hi] = 0
        sx[hi] = sn - 1
    a1 = np.clip(np.rint(fx * np.float32(2048)), -32768, 32767).astype(np.int32)
    a0 = np.clip(np.rint((np.float32(1.0) - fx) * np.float32(2048)), -32768, 32767).astype(np.int32)
    return np.stack([np.clip(sx, 0, sn - 1), np.clip(sx + 1, 0, sn - 1), a0, a1]).astype(np.int32)


@torch.no_grad()


@_lib.on_tensor_device
def prepare_images(bgr_u8: torch.Tensor, target_size: Tuple[int, int] = (640, 640), swap_rb: bool = True) -> torch.Tensor:
    """uint8 [N,Hs,Ws,3] (or [Hs,Ws,3]) CUDA tensor as ``cv2.imread`` lays it out -> float32 [N,3,H,W] in [0,1], RGB;
    ``target_size`` is (width, height) like the reference's."""
    if not bgr_u8.is_cuda:
        raise _lib.HisError("prepare_images: CUDA tensor required (no CPU fallback)")
    x = bgr_u8 if bgr_u8.dim() == 4 else bgr_u8[None]
    if x.dtype != torch.uint8 or x.shape[-1] != 3:
        raise ValueError("prepare_images expects uint8 [N,H,W,3]")
    x = x.contiguous()
    n, hs, ws, _ = x.shape
    wd, hd = int(target_size[0]), int(target_size[1])
    xt = torch.from_numpy(_tables(wd, ws, False)).to(x.device)
    yt = torch.from_numpy(_tables(hd, hs, True)).to(x.device)
    out = torch.empty((n, 3, hd, wd), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().his_preprocess_u8(x.data_ptr(), n, hs, ws, hd, wd, xt.data_ptr(), yt.data_ptr(), 1 if swap_rb else 0, out.data_ptr(),
                                             torch.cuda.current_stream(x.device).cuda_stream), "his_preprocess_u8")
    return out


def normalize_bbox(bbox: Sequence[float], img_width: int, img_height: int):
    """COCO [x,y,w,h] -> clipped normalised [x1,y1,x2,y2] (test_hierarchical_instance_peopleseg_onnx.py:118-141)."""
    x, y, w, h = bbox
    c = lambda v: max(0, min(1, v))          # noqa: E731
    return [c(x / img_width), c(y / img_height), c((x + w) / img_width), c((y + h) / img_height)]


def rois_from_boxes(boxes_per_image: Sequence[Sequence[Sequence[float]]], sizes: Sequence[Tuple[int, int]]) -> torch.Tensor:
    """[[COCO boxes of image 0], [of image 1], ...] + [(width, height), ...] -> float32 [K,5] = [batch_idx, x1, y1, x2, y2]."""
    rows = []
    for b, (boxes, (w, h)) in enumerate(zip(boxes_per_image, sizes)):
        for bb in boxes:
            rows.append([float(b)] + [float(v) for v in normalize_bbox(bb, w, h)])
    return torch.tensor(rows, dtype=torch.float32).reshape(-1, 5)
