"""Oracle (test infrastructure): host-side pre-processing of the demo / validation input path
(test_hierarchical_instance_peopleseg_onnx.py:118-141,170-196).

``prepare_image`` there is cv2.imread -> BGR2RGB -> cv2.resize(INTER_LINEAR, uint8) -> /255 -> CHW.  The arithmetic that matters
is OpenCV's 8-bit bilinear resize (third-party, opencv-python 4.13 in this image), restated here in integer numpy:
coefficients in 11-bit fixed point from a float ``fx`` whose scale is ``1/(dst/src)`` in double, horizontal pass in int32,
vertical pass ``(((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2``.  Pinned against cv2.resize itself (tests/test_preprocess.py)."""
from __future__ import annotations

import numpy as np


def linear_coefs(dn: int, sn: int, vertical: bool = False):
    """Source indices and 11-bit weights of cv2's INTER_LINEAR.  Horizontally the weight is forced to (1, 0) where the tap
    leaves the image; vertically cv2 keeps the fractional weights and only clips the two ROW INDICES (so a border row is
    blended with itself through two separately truncated products)."""
    scale = 1.0 / (float(dn) / float(sn))                   # cv2: scale_x = 1./inv_scale_x, both double
    d = np.arange(dn)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    fx = (fx - sx).astype(np.float32)
    if not vertical:
        lo = sx < 0
        fx[lo] = 0; sx[lo] = 0
        hi = sx >= sn - 1
        fx[hi] = 0; sx[hi] = sn - 1
    a1 = np.clip(np.rint(fx * np.float32(2048)), -32768, 32767).astype(np.int32)
    a0 = np.clip(np.rint((np.float32(1.0) - fx) * np.float32(2048)), -32768, 32767).astype(np.int32)
    return np.clip(sx, 0, sn - 1), np.clip(sx + 1, 0, sn - 1), a0, a1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh)) for uint8 HxWxC, INTER_LINEAR."""
    sh, sw = src.shape[:2]
    x0, x1, ax0, ax1 = linear_coefs(dw, sw)
    y0, y1, ay0, ay1 = linear_coefs(dh, sh, vertical=True)
    s = src.astype(np.int32)
    rows = s[:, x0] * ax0[None, :, None] + s[:, x1] * ax1[None, :, None]
    r0, r1 = rows[y0], rows[y1]
    out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def prepare_image(bgr: np.ndarray, target_size=(640, 640)) -> np.ndarray:
    """prepare_image (test_hierarchical_instance_peopleseg_onnx.py:170-196) after the file decode: [1,3,H,W] float32 RGB in [0,1]."""
    rgb = bgr[:, :, ::-1]
    r = resize_linear_u8(rgb, target_size[0], target_size[1])
    return np.transpose(r.astype(np.float32) / np.float32(255.0), (2, 0, 1))[None]


def normalize_bbox(bbox, img_width, img_height):
    """normalize_bbox (:118-141): COCO [x,y,w,h] -> clipped normalised [x1,y1,x2,y2] (Python float arithmetic)."""
    x, y, w, h = bbox
    c = lambda v: max(0, min(1, v))
    return [c(x / img_width), c(y / img_height), c((x + w) / img_width), c((y + h) / img_height)]
