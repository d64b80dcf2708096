"""the input path (prepare_image / normalize_bbox of test_hierarchical_instance_peopleseg_onnx.py:118-196).
CPU: the oracle's integer restatement of OpenCV's 8-bit bilinear resize against cv2.resize itself (cv2 is the third-party code the
reference calls; it is importable here and on the GPU box) and normalize_bbox against the reference function.  GPU: the fused
kernel against cv2, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import preport, refload

SIZES = [(480, 640, 480, 640), (375, 500, 480, 640), (720, 1280, 480, 640), (333, 517, 120, 160), (100, 100, 640, 640),
         (427, 640, 640, 640), (31, 47, 64, 48), (3, 5, 64, 48), (97, 13, 33, 200)]


def _img(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def test_resize_port_is_cv2_bit_exact():
    import cv2
    for i, (sh, sw, dh, dw) in enumerate(SIZES):
        src = _img(i, sh, sw)
        assert np.array_equal(preport.resize_linear_u8(src, dw, dh), cv2.resize(src, (dw, dh))), (sh, sw, dh, dw)


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present (GPU box)")
def test_normalize_bbox_matches_reference_function():
    from human_instance_segmentation_b200 import preprocess
    fn = refload.ref_functions_from_script("test_hierarchical_instance_peopleseg_onnx.py", ["normalize_bbox"])["normalize_bbox"]
    for bb, w, h in [([10.5, 20.25, 300, 200], 640, 480), ([-5, 3, 700, 500], 640, 480), ([0, 0, 1, 1], 3, 7), ([600, 470, 100, 100], 640, 480)]:
        assert preprocess.normalize_bbox(bb, w, h) == fn(bb, w, h) == preport.normalize_bbox(bb, w, h)
    rois = preprocess.rois_from_boxes([[[10, 20, 100, 50]], [], [[0, 0, 640, 480], [320, 240, 1000, 10]]], [(640, 480)] * 3)
    assert rois.shape == (3, 5) and rois[:, 0].tolist() == [0.0, 2.0, 2.0] and float(rois[2, 3]) == 1.0


@pytest.mark.gpu
def test_fused_preprocess_kernel_matches_cv2_pipeline():
    import cv2
    from human_instance_segmentation_b200 import preprocess
    for i, (sh, sw, dh, dw) in enumerate(SIZES):
        batch = np.stack([_img(10 * i + k, sh, sw) for k in range(2)])
        got = preprocess.prepare_images(torch.from_numpy(batch).cuda(), (dw, dh)).cpu().numpy()
        for k in range(2):
            rgb = cv2.cvtColor(batch[k], cv2.COLOR_BGR2RGB)                       # prepare_image, :182-195
            want = np.transpose(cv2.resize(rgb, (dw, dh)).astype(np.float32) / 255.0, (2, 0, 1))
            assert np.array_equal(got[k], want), (sh, sw, dh, dw)
    one = preprocess.prepare_images(torch.from_numpy(_img(99, 50, 70)).cuda(), (32, 24), swap_rb=False)
    assert one.shape == (1, 3, 24, 32) and np.array_equal(one[0].cpu().numpy(), preport.prepare_image(_img(99, 50, 70)[:, :, ::-1], (32, 24))[0])
