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
