"""GPU: post-processing stencil kernels (through the reference-shaped modules of the product package) against the golden
vectors of the reference modules.  Integer stages (argmax -> instance mask, NEAREST paste-back) are bit-exact.  The
float stencils threshold a value that can sit within a few ulp of the threshold: they must be exact
on every pixel whose reference margin |v - thr| exceeds 1e-5, and the count of differing pixels is bounded and printed."""
import numpy as np
import pytest
import torch

from human_instance_segmentation_b200 import postprocess as pp
from oracle import postport
from tests import common

pytestmark = pytest.mark.gpu


def _mismatch(got, want):
    return int((got.cpu() != want).sum())


def test_edge_smoothing_matches_reference():
    g = common.golden("post")
    masks = g["masks"].cuda()
    assert _mismatch(pp.BinaryMaskEdgeSmoothing()(masks), g["edge_smooth"]) == 0
    assert _mismatch(pp.BinaryMaskEdgeSmoothing(0.4, 2.0)(masks), g["edge_smooth_t04_s2"]) == 0
    # all 512 binary 3x3 neighbourhoods, including the exact-tie patterns
    assert _mismatch(pp.BinaryMaskEdgeSmoothing()(g["patterns"].cuda()), g["edge_smooth_patterns"]) == 0
    # [H,W] and [C,H,W] inputs keep their shape (edge_smoothing.py:45-50,84-88)
    assert pp.BinaryMaskEdgeSmoothing()(masks[0, 0]).shape == (96, 128)
    assert pp.BinaryMaskEdgeSmoothing()(masks[0]).shape == (1, 96, 128)


def _check_soft(got, want, soft, thr, what, max_bad=0):
    bad = (got.cpu() != want)
    safe_bad = int((bad & ((soft - thr).abs() > 1e-5)).sum())
    print(f"{what}: {int(bad.sum())} differing pixels of {want.numel()}, {safe_bad} outside the 1e-5 tie band")
    assert safe_bad == 0 and int(bad.sum()) <= max_bad


def test_bilateral_filters_match_reference():
    g = common.golden("post")
    for key, x in (("", g["masks"]), ("_noisy", g["noisy"])):
        _check_soft(pp.BinaryMaskBilateralFilter()(x.cuda()), g["binary_bilateral" + key], postport.binary_bilateral(x, return_soft=True), 0.5,
                    "BinaryMaskBilateralFilter" + key, max_bad=2)
        _check_soft(pp.MorphologicalBilateralFilter()(x.cuda()), g["morph_bilateral" + key], postport.morph_bilateral(x, return_soft=True), 0.5,
                    "MorphologicalBilateralFilter" + key, max_bad=2)
    _check_soft(pp.BinaryMaskBilateralFilter(5, 1.0, 0.5, 3)(g["masks"].cuda()), g["binary_bilateral_k5_it3"],
                postport.binary_bilateral(g["masks"], 5, 1.0, 0.5, 3, return_soft=True), 0.5, "BinaryMaskBilateralFilter(k5,it3)", max_bad=2)


def test_instance_mask_and_paste_back_bit_exact():
    g = common.golden("post")
    logits, rois = g["paste_logits"].cuda(), g["paste_rois"].cuda()
    target = pp.instance_masks(logits, score_threshold=0.5, as_uint8=True)
    assert np.array_equal(target.cpu().numpy(), g["paste_target"].numpy())
    canvas = pp.paste_masks(target, rois, 1, 480, 640)
    assert np.array_equal(canvas.cpu().numpy(), g["paste_canvas"].numpy())
    plain = pp.instance_masks(logits)
    assert torch.equal(plain.cpu(), postport.instance_mask(g["paste_logits"]))
    # ties: argmax returns the FIRST maximum -> class 1 needs l1 > l0 and l1 >= l2
    t = torch.tensor([[1.0, 1.0, 1.0], [0.0, 1.0, 1.0], [0.0, 1.0, 2.0], [2.0, 2.0, 0.0]]).t().reshape(1, 3, 1, 4).contiguous()
    assert pp.instance_masks(t.cuda()).flatten().tolist() == [0.0, 1.0, 0.0, 0.0]


def test_paste_back_full_size_property():
    """BASELINE config-5 geometry (640x480 canvas, 128x96 masks) on fresh seeded boxes, against the CPU oracle."""
    g = torch.Generator().manual_seed(33)
    n = 64
    masks = (torch.rand(n, 128, 96, generator=g) > 0.4).to(torch.uint8)
    rois = common.synth_rois(33, 8, 8)
    canvas = pp.paste_masks(masks.cuda(), rois.cuda(), 8, 480, 640).cpu().numpy()
    assert np.array_equal(canvas, postport.paste_back(masks.numpy(), rois.numpy(), 8, 480, 640))
    assert pp.paste_masks(masks[:0].cuda(), rois[:0].cuda(), 2, 48, 64).sum() == 0


def test_mask_dilation_matches_reference_golden():
    g = common.golden("small_b0_bn_relu")
    logits = g["logits"].cuda()
    for d in (1, 2):
        out = pp.MaskDilationModule(d)(logits).cpu()
        want = g[f"dilated{d}"]
        bad = int((out != want).sum())
        # softmax on the GPU differs from torch CPU by ~1 ulp; the 0.1 threshold on (dilated - p) may flip on exact ties only
        margin = ((torch.nn.functional.max_pool2d(torch.softmax(g["logits"], 1)[:, 1:2], 2 * d + 1, 1, d) - torch.softmax(g["logits"], 1)[:, 1:2]) - 0.1).abs()
        assert bad == 0 or float(margin.min()) < 1e-6, bad
    assert torch.equal(pp.MaskDilationModule(0)(logits), logits)
