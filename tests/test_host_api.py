"""CPU: host-side logic of the product package (no compute): state-dict contract, factory behaviour, C-ABI exports."""
import ctypes
import os
import re

import pytest
import torch

import human_instance_segmentation_b200 as his
from oracle import headport
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("preset", ["b0", "b1_enhanced", "b7_ultra"])
def test_state_dict_keys_and_shapes_match_reference(preset):
    cfg = headport.PRESETS[preset]
    model = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
    want = common.golden_keys()["preset_" + preset]
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert list(got.keys()) == list(want.keys())
    assert got == want


@pytest.mark.parametrize("name", list(common.GUIDED_CASES))
def test_guided_variant_state_dict_matches_reference(name):
    cfg = common.GUIDED_CASES[name][0]
    model = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
    want = common.golden_keys()[name]
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want and list(got) == list(want)
    assert model.segmentation_head.final_classifier[-1].bias.tolist() == [0.0, 0.0, -0.5]      # rgb.py:117-123
    assert not hasattr(model, "feature_combiner")


@pytest.mark.parametrize("name", list(common.REFINE_CASES))
def test_refinement_flag_state_dict_matches_reference(name):
    cfg = common.REFINE_CASES[name][0]
    model = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
    want = common.golden_keys()[name]
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want and list(got) == list(want)
    if cfg.use_boundary_refinement:
        assert float(model.segmentation_head.boundary_refiner.blend_weight.detach()) == pytest.approx(0.01)


@pytest.mark.parametrize("name", list(common.MULTISCALE_CASES))
def test_multiscale_state_dict_matches_reference(name):
    cfg = common.MULTISCALE_CASES[name][0]
    model = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
    want = common.golden_keys()[name]
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want and list(got) == list(want)
    assert model.scales == [n for n, _ in cfg.roi_sizes] and all(not ra.aligned for ra in model.roi_aligns.values())     # rgb.py:826-833


@pytest.mark.parametrize("name", list(common.STANDARD_CASES))
def test_standard_variant_state_dict_matches_reference(name):
    cfg = common.STANDARD_CASES[name][0]
    model = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
    want = common.golden_keys()[name]
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want and list(got) == list(want)
    assert model.roi_align.aligned is False and model.roi_align.spatial_scale == 640.0          # rgb.py:404-408


def test_roi_level_unet_variant_is_unconstructible_like_the_reference():
    kw = dict(headport.PRESETS["b0"].factory_kwargs(), use_full_image_unet=False)
    with pytest.raises(NotImplementedError):     # the reference raises NameError at rgb.py:497
        his.create_rgb_hierarchical_model(**kw)


def test_reference_checkpoint_roundtrip_and_pinned_constants():
    cfg = headport.PRESETS["b0"]
    model = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
    w = model.pretrained_unet.output_conv.weight.flatten().tolist()
    assert w == [1.0, -1.0] and model.pretrained_unet.output_conv.bias.abs().sum() == 0   # ..._unet.py:1963-1971
    assert model.pretrained_unet.model.mean == [0.485, 0.456, 0.406]                        # path string contains "b0"
    sd = common.procedural_state(common.golden_keys()["preset_b0"])
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert model.roi_align_mask.spatial_scale == 640.0 and model.roi_align_rgb.aligned is True  # rgb.py:636-647


def test_normalisation_follows_path_string():
    m = his.PreTrainedPeopleSegmentationUNet(pretrained_weights_path="ext_extractor/2020-09-23a.pth", encoder_name="timm-efficientnet-b3")
    assert m.mean == [0.5, 0.5, 0.5] and m.std == [0.5, 0.5, 0.5]


def test_factory_errors_match_reference_behaviour():
    kw = headport.PRESETS["b0"].factory_kwargs()
    with pytest.raises(ValueError):     # activation_utils.py:103
        his.create_rgb_hierarchical_model(**{**kw, "activation_function": "tanh"})
    with pytest.raises(ValueError):     # normalization_comparison.py:206
        his.create_rgb_hierarchical_model(**{**kw, "normalization_type": "nonsense"})


def test_no_cpu_fallback():
    model = his.create_rgb_hierarchical_model(**headport.PRESETS["b0"].factory_kwargs())
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(his.HisError):
        model(torch.rand(1, 3, 64, 64), torch.tensor([[0, 0.1, 0.1, 0.9, 0.9]]))
    with pytest.raises(his.HisError):
        his.postprocess.instance_masks(torch.zeros(1, 3, 4, 4))


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "his_b200.h")).read()
    names = sorted(set(re.findall(r"\b(his_[A-Za-z0-9_]+)\s*\(", header)))
    assert len(names) >= 30
    path = his.build()
    lib = ctypes.CDLL(path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/his_b200.h but not exported"
    from human_instance_segmentation_b200 import lib as L
    assert set(L.SIGNATURES) == set(names)
    assert lib.his_version() == his.lib.ABI_VERSION


def test_tile_n_policy():
    lib = his.load()
    nt, bn = ctypes.c_int(), ctypes.c_int()
    for cout, want in [(256, (1, 256)), (128, (1, 128)), (72, (1, 80)), (36, (1, 48)), (288, (2, 160)), (384, (2, 192)), (768, (3, 256)), (16, (1, 16))]:
        assert lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn)) == 0
        assert (nt.value, bn.value) == want, cout


def test_pretrained_unet_weights_are_loaded_from_the_path(tmp_path):
    """..._unet.py:1776-1864: the constructor loads `pretrained_weights_path` when the file exists (plain or wrapped state
    dict, optional 'model.' / 'unet.' prefix, strict=False) and warns when it does not."""
    import warnings
    from human_instance_segmentation_b200 import param_tree as pt
    donor = pt.SmpUnetParams("timm-efficientnet-b0")
    g = torch.Generator().manual_seed(5)
    sd = {k: (torch.randn(v.shape, generator=g) if v.dtype.is_floating_point else v.clone()) for k, v in donor.state_dict().items()}
    for wrap, prefix in (("state_dict", "model."), ("model_state_dict", "unet."), (None, "")):
        f = tmp_path / f"best_model_b0_{wrap}.pth"
        payload = {prefix + k: v for k, v in sd.items()}
        torch.save({wrap: payload, "epoch": 3} if wrap else payload, f)
        with warnings.catch_warnings():
            warnings.simplefilter("error")          # a complete file loads silently
            m = his.create_rgb_hierarchical_model(roi_size=(16, 12), mask_size=(32, 24), use_pretrained_unet=True, use_full_image_unet=True,
                                                  pretrained_weights_path=str(f), encoder_name="timm-efficientnet-b0", use_contour_detection=True,
                                                  normalization_type="batchnorm")
        inner = m.pretrained_unet.model
        assert inner.load_report == ([], [])
        for k, v in inner.model.state_dict().items():
            assert torch.equal(v, sd[k]), k
        assert inner.mean == [0.485, 0.456, 0.406]                     # "b0" in the path string (:1744-1758)
    with pytest.warns(UserWarning, match="not found"):
        his.create_rgb_hierarchical_model(use_pretrained_unet=True, use_full_image_unet=True, pretrained_weights_path=str(tmp_path / "absent.pth"),
                                          encoder_name="timm-efficientnet-b0")
    # a partial file: strict=False semantics, reported
    part = {k: v for i, (k, v) in enumerate(sd.items()) if i % 2 == 0}
    torch.save(part, tmp_path / "half_b0.pth")
    with pytest.warns(UserWarning, match="missing"):
        m = his.create_rgb_hierarchical_model(use_pretrained_unet=True, use_full_image_unet=True, pretrained_weights_path=str(tmp_path / "half_b0.pth"),
                                              encoder_name="timm-efficientnet-b0")
    # (torch does not report BatchNorm's num_batches_tracked as missing for a state dict without version metadata)
    assert set(m.pretrained_unet.model.load_report[0]) == {k for k in sd if k not in part and not k.endswith("num_batches_tracked")}


def test_plan_capacity_buckets_and_split_weight_packing():
    """Host logic of the serving plan cache (dynamic batch_size / num_rois, export_onnx_advanced.py:427-457) and of the strict
    precision mode's weight split."""
    from human_instance_segmentation_b200 import engine, model
    prev = 0
    for n in range(0, 3000):
        b = model._bucket_rois(n)
        assert b >= n and b >= prev and model._bucket_rois(b) == b          # covers the request, monotonic, idempotent
        assert b - n < (1 if n <= 16 else 8 if n <= 64 else 32 if n <= 256 else 64)     # bounded padding
        prev = b
    assert [model._bucket_images(b) for b in (0, 1, 16, 17, 63, 64)] == [0, 1, 16, 24, 64, 64]
    g = torch.Generator().manual_seed(0)
    w = torch.randn(40, 24, 3, 3, generator=g) * 0.07
    scale = torch.rand(40, generator=g) + 0.5
    hi, k1 = engine.pack_gemm_weight(w, 48, False, scale)
    both, k2 = engine.pack_gemm_weight(w, 48, False, scale, split=True)
    assert k2 == 2 * k1 and both.shape[-1] == k2 and torch.equal(both[..., :k1], hi)
    full = (w * scale.view(-1, 1, 1, 1)).permute(2, 3, 0, 1).reshape(9, 40, 24)
    rec = both[0, :, :40, :24].float() + both[0, :, :40, k1:k1 + 24].float()
    assert float((rec - full).abs().max()) <= 2.0 ** -21 * float(full.abs().max())      # hi + lo carries ~21 bits
    assert float((hi[0, :, :40, :24].float() - full).abs().max()) > 2.0 ** -14 * float(full.abs().max())   # ... a single fp16 does not
    wt = torch.randn(24, 40, 2, 2, generator=g)
    bt, kt = engine.pack_gemm_weight(wt, 48, True, None, split=True)
    assert bt.shape == (4, 1, 48, kt) and kt == 128


def test_bench_reference_arm_prints_exactly_one_json_line():
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == "roi_masks_per_sec" and rec["value"] > 0
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["cpu_baseline"]["kind"] == "port"
