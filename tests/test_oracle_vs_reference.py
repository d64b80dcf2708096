"""CPU, authoring container only: the oracle port against the reference modules run live."""
import pytest
import torch

from oracle import headport, paramfill, refload
from tests import common

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present (GPU box)")


def test_state_dict_keys_fixture_is_current():
    cfg = headport.PRESETS["b0"]
    model = refload.build_reference_model(**cfg.factory_kwargs())
    want = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert common.golden_keys()["preset_b0"] == want


def test_live_reference_equals_port_nonpreset_flags():
    from dataclasses import replace
    cfg = replace(headport.PRESETS["b0"], roi_size=(8, 12), mask_size=(16, 24), use_contour_detection=False,
                  activation_function="gelu")
    model = refload.build_reference_model(**cfg.factory_kwargs())
    sd = paramfill.fill_state_dict(model.state_dict(), seed=5)
    model.load_state_dict(sd)
    images = common.synth_images(3, 1, 64, 64)
    rois = common.synth_rois(3, 1, 3)
    with torch.no_grad():
        ref, ref_aux = model(images, rois)
    out, aux = headport.forward(sd, images, rois, cfg)
    assert common.rel_err(out, ref) < 2e-5
    assert "contours" not in aux and "contours" not in ref_aux
