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


@pytest.mark.parametrize("d", [1, 2, 5, 16])
def test_fused_dilation_instance_mask_equals_the_two_steps(d):
    """instance_masks(..., dilation_pixels=d) == instance_masks(MaskDilationModule(d)(logits)) bit for bit (ROI-sized planes,
    ragged planes with partial tiles, score threshold on / off, fp32 and u8 outputs), and the tiled dilation equals a
    max_pool2d restatement of the module on the same softmax values."""
    g = torch.Generator().manual_seed(7 + d)
    for shape in ((5, 3, 128, 96), (2, 3, 37, 70), (1, 3, 3, 2)):
        lg = (torch.randn(shape[0], 3, max(shape[2] // 8, 1), max(shape[3] // 8, 1), generator=g) * 2)
        lg = torch.nn.functional.interpolate(lg, size=shape[2:], mode="bilinear").contiguous().cuda()
        dl = pp.MaskDilationModule(d)(lg)
        for thr in (0.0, 0.5):
            assert torch.equal(pp.instance_masks(lg, thr, dilation_pixels=d), pp.instance_masks(dl, thr))
            assert torch.equal(pp.instance_masks(lg, thr, as_uint8=True, dilation_pixels=d), pp.instance_masks(dl, thr, as_uint8=True))
        # the module restated with torch ops on the GPU's own fp32 softmax: identical up to ties of (dilated - p) with 0.1
        p1 = torch.softmax(lg, 1)[:, 1:2]
        grow = (torch.nn.functional.max_pool2d(p1, 2 * d + 1, 1, d) - p1)
        want = lg.clone(); want[:, 1:2] += 2.0 * (grow > 0.1).float()
        bad = (dl != want)
        assert int(bad.sum()) == 0 or float((grow - 0.1).abs()[bad[:, 1:2]].max()) < 1e-6


def _frac_bad(got, want):
    return float((got.float().cpu() != want).float().mean())


def test_edge_smoothing_variants_match_reference():
    """a16 variants (export_edge_smoothing_onnx.py:63-318, hed/edge_smoothing.py:93-170) against the reference's outputs."""
    g = common.golden("post_variants")
    m = g["masks"].cuda()
    # linear-arithmetic variants: bit-exact, including every binary 3x3 neighbourhood
    assert _mismatch(pp.AdaptiveEdgeSmoothing()(m, g["ad_bs"].cuda(), g["ad_sens"].cuda(), g["ad_thr"].cuda()), g["adaptive"]) == 0
    assert _mismatch(pp.OptimizedEdgeSmoothing(use_fp16=False)(m), g["optimized_fp32"]) == 0
    assert _mismatch(pp.OptimizedEdgeSmoothing(use_fp16=False)(g["patterns"].cuda()), g["optimized_fp32_patterns"]) == 0
    # atan2/cos/sin (directional) and half-precision rounding order (fp16 graph) differ from the CPU libraries by ulps:
    # only pixels within a hair of the 0.5 threshold may flip
    bad = _frac_bad(pp.DirectionalEdgeSmoothing()(m), g["directional"])
    print(f"DirectionalEdgeSmoothing: {bad:.2e} of pixels differ")
    assert bad <= 2e-4
    out16 = pp.OptimizedEdgeSmoothing(use_fp16=True)(m)
    assert out16.dtype == torch.float16
    bad = _frac_bad(out16, g["optimized_fp16"])
    print(f"OptimizedEdgeSmoothing(fp16): {bad:.2e} of pixels differ")
    assert bad <= 1e-3
    mc = pp.MultiClassEdgeSmoothing()
    assert _mismatch(mc.smooth_predictions(g["logits"].cuda()), g["multiclass3"]) == 0
    assert _mismatch(pp.MultiClassEdgeSmoothing(0.5, 3.0, 2).smooth_predictions(g["logits"].cuda(), apply_softmax=True), g["multiclass3_softmax_it2"]) == 0
    assert _mismatch(mc.smooth_predictions(g["probs5"].cuda()), g["multiclass5"]) == 0
    assert mc.smooth_predictions(g["logits"][0].cuda()).shape == (3, 48, 64)


def test_value_filters_match_reference():
    """a17: BilateralFilter / FastBilateralFilter / EdgePreservingFilter return float images -> absolute tolerance 2e-5
    (expf and summation order differ from the CPU by a few ulp)."""
    g = common.golden("post_variants")
    gray, guide = g["gray"].cuda(), g["guide"].cuda()
    tol = 2e-5
    assert (pp.BilateralFilter()(gray[:, :, :12, :16].contiguous()).cpu() - g["bilateral_exact"]).abs().max() < tol
    assert (pp.BilateralFilter(3, 0.8, 0.3)(gray[:1, :, :10, :12].contiguous()).cpu() - g["bilateral_exact_k3"]).abs().max() < tol
    assert (pp.FastBilateralFilter()(gray).cpu() - g["bilateral_fast"]).abs().max() < tol
    assert (pp.FastBilateralFilter(7, 1.5, 0.2, 3)(gray).cpu() - g["bilateral_fast_k7_it3"]).abs().max() < tol
    assert (pp.EdgePreservingFilter()(gray).cpu() - g["edge_preserving"]).abs().max() < 1e-4
    assert (pp.EdgePreservingFilter(3, 0.05)(gray, guide).cpu() - g["edge_preserving_guided_r3"]).abs().max() < 1e-4
    # full-size image against the oracle (the reference's Python triple loop cannot run at this size)
    big = torch.rand(1, 1, 120, 160, generator=torch.Generator().manual_seed(4))
    assert (pp.BilateralFilter(7, 1.5, 0.2)(big.cuda()).cpu() - postport.exact_bilateral(big, 7, 1.5, 0.2)).abs().max() < tol
    with pytest.raises(ValueError):
        pp.BilateralFilter(4)


def test_fused_mask_cleanup_equals_the_two_modules_and_the_oracle():
    """BASELINE config 5 chain on 480x640 masks: one fused shared-memory pass == edge smoothing then bilateral filter."""
    from oracle.make_golden_post import blob_masks
    masks = blob_masks(5, 6, 480, 640)
    fused = pp.MaskCleanup()(masks.cuda())
    two = pp.BinaryMaskBilateralFilter()(pp.BinaryMaskEdgeSmoothing()(masks.cuda()))
    assert torch.equal(fused, two)
    want_soft = postport.binary_bilateral(postport.edge_smooth(masks), return_soft=True)
    _check_soft(fused, (want_soft > 0.5).float(), want_soft, 0.5, "MaskCleanup 480x640", max_bad=4)
    # ragged sizes (partial tiles on both axes), several planes per image
    odd = blob_masks(9, 2, 70, 45).repeat(1, 3, 1, 1).contiguous()
    assert torch.equal(pp.MaskCleanup(num_iterations=1, kernel_size=5)(odd.cuda()),
                       pp.BinaryMaskBilateralFilter(5, 1.5, 0.5, 1)(pp.BinaryMaskEdgeSmoothing()(odd.cuda())))
    want = postport.binary_bilateral(postport.edge_smooth(odd), 5, 1.5, 0.5, 1)
    assert _frac_bad(pp.MaskCleanup(num_iterations=1, kernel_size=5)(odd.cuda()), want) <= 1e-4


@pytest.mark.parametrize("k,its", [(7, 2), (7, 1), (5, 2), (5, 1), (3, 2), (3, 1)])
def test_wide_mask_cleanup_is_bit_identical_to_the_unfused_chain(k, its):
    """The 64x32 / vector-load / table form of the fused clean-up kernel (post_stencil.cu: mask_cleanup_wide_kernel) against
    the two separate kernels that keep the plain per-pixel arithmetic: noise masks (every (centre, edge, corner) neighbour
    count of the edge-smoothing table occurs), blob masks, soft masks (the block-wide vote sends them down the general
    edge-smoothing path), sizes with partial tiles on both axes and W % 4 != 0 (scalar stores)."""
    from oracle.make_golden_post import blob_masks
    g = torch.Generator().manual_seed(100 * k + its)
    fused = pp.MaskCleanup(kernel_size=k, num_iterations=its).cuda()
    es, bf = pp.BinaryMaskEdgeSmoothing().cuda(), pp.BinaryMaskBilateralFilter(k, 1.5, 0.5, its).cuda()
    cases = [
        (torch.rand(3, 1, 96, 128, generator=g) > 0.5).float(),
        (torch.rand(2, 2, 67, 131, generator=g) > 0.3).float(),          # W % 4 != 0, partial tiles
        (torch.rand(1, 1, 33, 65, generator=g) > 0.7).float(),
        (torch.rand(1, 1, 5, 3, generator=g) > 0.5).float(),             # smaller than the halo
        blob_masks(3, 2, 120, 200),
        torch.rand(2, 1, 70, 90, generator=g),                           # soft masks
        torch.cat([(torch.rand(1, 1, 64, 128, generator=g) > 0.5).float(), torch.rand(1, 1, 64, 128, generator=g)]),   # mixed planes
    ]
    for m in cases:
        m = m.cuda().contiguous()
        want = bf(es(m))
        got = fused(m)
        assert torch.equal(got, want), f"k={k} its={its} shape={tuple(m.shape)}: {(got != want).sum().item()} pixels differ"
    # an output buffer that is not 16-byte aligned (view at an odd element offset): scalar stores
    m = (torch.rand(2, 1, 64, 128, generator=g) > 0.5).float().cuda()
    buf = torch.empty(m.numel() + 1, device="cuda")
    out = buf[1:].view_as(m)
    assert torch.equal(fused(m, out=out), bf(es(m)))
    # byte masks: same kernel, 0 / 1 in and out (aligned and odd widths, a view at an odd byte offset)
    for shape in ((3, 1, 96, 128), (2, 1, 67, 131)):
        m = (torch.rand(*shape, generator=g) > 0.5)
        assert torch.equal(fused(m.to(torch.uint8).cuda()), bf(es(m.float().cuda())).to(torch.uint8))
    m = (torch.rand(2, 1, 64, 128, generator=g) > 0.5)
    buf8 = torch.empty(m.numel() + 1, dtype=torch.uint8, device="cuda")
    assert torch.equal(fused(m.to(torch.uint8).cuda(), out=buf8[1:].view(2, 1, 64, 128)), bf(es(m.float().cuda())).to(torch.uint8))
    # other thresholds / strength
    f2 = pp.MaskCleanup(0.4, 2.0, k, 1.2, 0.45, its).cuda()
    m = blob_masks(4, 2, 90, 150).cuda()
    assert torch.equal(f2(m), pp.BinaryMaskBilateralFilter(k, 1.2, 0.45, its).cuda()(pp.BinaryMaskEdgeSmoothing(0.4, 2.0).cuda()(m)))
