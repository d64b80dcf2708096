"""CPU: oracle/postport.py against the golden vectors produced by the reference post-processing modules/functions."""
import numpy as np
import torch

from oracle import postport
from tests import common


def test_postport_matches_reference_goldens():
    g = common.golden("post")
    masks, noisy = g["masks"], g["noisy"]
    assert torch.equal(postport.edge_smooth(masks), g["edge_smooth"])
    assert torch.equal(postport.edge_smooth(masks, 0.4, 2.0), g["edge_smooth_t04_s2"])
    assert torch.equal(postport.edge_smooth(g["patterns"]), g["edge_smooth_patterns"])
    assert torch.equal(postport.binary_bilateral(masks), g["binary_bilateral"])
    assert torch.equal(postport.binary_bilateral(noisy), g["binary_bilateral_noisy"])
    assert torch.equal(postport.binary_bilateral(masks, 5, 1.0, 0.5, 3), g["binary_bilateral_k5_it3"])
    assert torch.equal(postport.morph_bilateral(masks), g["morph_bilateral"])
    assert torch.equal(postport.morph_bilateral(noisy), g["morph_bilateral_noisy"])


def test_paste_back_port_matches_reference_script():
    g = common.golden("post")
    target = postport.instance_mask(g["paste_logits"], 0.5)[:, 0].numpy().astype(np.uint8)
    assert np.array_equal(target, g["paste_target"].numpy())
    canvas = postport.paste_back(target, g["paste_rois"].numpy(), 1, 480, 640)
    assert np.array_equal(canvas, g["paste_canvas"].numpy())


def test_nearest_index_is_cv2_rule():
    import cv2
    for src in (24, 32, 96, 128):
        for dst in (1, 7, 50, 97, 200, 333):
            ramp = np.arange(src, dtype=np.uint8)[None, :].repeat(2, 0)
            ref = cv2.resize(ramp, (dst, 2), interpolation=cv2.INTER_NEAREST)[0]
            assert np.array_equal(ref, postport.nearest_index(dst, src).astype(np.uint8)), (src, dst)


def test_post_variant_ports_match_reference_goldens():
    """a16/a17 variants (export_edge_smoothing_onnx.py:63-318, hed/edge_smoothing.py:93-170, hed/bilateral_filter.py:9-296)."""
    g = common.golden("post_variants")
    m = g["masks"]
    assert torch.equal(postport.directional_edge_smooth(m), g["directional"])
    assert torch.equal(postport.adaptive_edge_smooth(m, g["ad_bs"], g["ad_sens"], g["ad_thr"]), g["adaptive"])
    assert torch.equal(postport.optimized_edge_smooth(m), g["optimized_fp32"])
    assert torch.equal(postport.optimized_edge_smooth(g["patterns"]), g["optimized_fp32_patterns"])
    assert torch.equal(postport.multiclass_edge_smooth(g["logits"]), g["multiclass3"])
    assert torch.equal(postport.multiclass_edge_smooth(g["logits"], iterations=2, apply_softmax=True), g["multiclass3_softmax_it2"])
    assert torch.equal(postport.multiclass_edge_smooth(g["probs5"]), g["multiclass5"])
    gray, guide = g["gray"], g["guide"]
    assert (postport.exact_bilateral(gray[:, :, :12, :16]) - g["bilateral_exact"]).abs().max() < 1e-6
    assert (postport.exact_bilateral(gray[:1, :, :10, :12], 3, 0.8, 0.3) - g["bilateral_exact_k3"]).abs().max() < 1e-6
    assert (postport.fast_bilateral(gray) - g["bilateral_fast"]).abs().max() < 1e-6
    assert (postport.fast_bilateral(gray, 7, 1.5, 0.2, 3) - g["bilateral_fast_k7_it3"]).abs().max() < 1e-6
    assert (postport.edge_preserving(gray) - g["edge_preserving"]).abs().max() < 1e-5
    assert (postport.edge_preserving(gray, guide, 3, 0.05) - g["edge_preserving_guided_r3"]).abs().max() < 1e-5
    # the fp16 flavour of OptimizedEdgeSmoothing only differs from the fp32 one on near-tie pixels
    assert float((g["optimized_fp16"] != g["optimized_fp32"]).float().mean()) < 2e-3
