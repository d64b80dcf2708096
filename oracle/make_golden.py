"""Oracle (test infrastructure): generate tests/golden/* by running the REAL reference.

Run in the authoring container only (needs /root/reference):

    python -m oracle.make_golden [--only model|roi|post]

Everything stored is an *output of reference code* (with the smp.Unet stub of
oracle/refload.py standing in for the absent third-party UNet) on seeded inputs and the
procedural weights of oracle/paramfill.py.  The tests regenerate the same inputs and
weights and compare.
"""
from __future__ import annotations

import argparse
import json
import os
from dataclasses import replace

import numpy as np
import torch

from . import headport, paramfill, refload

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


from human_instance_segmentation_b200.synthetic import edge_rois, synth_images, synth_rois  # noqa: E402,F401


SMALL_CASES = {
    # name: (cfg, image hw, rois)
    "small_b0_bn_relu": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24)), (96, 128)),
    "small_b0_bn_relu_exportscale": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24),
                                             spatial_scale=(96.0, 128.0)), (96, 128)),
    "small_b1_bc72": (replace(headport.PRESETS["b1_enhanced"], roi_size=(20, 16), mask_size=(40, 32)), (64, 96)),
    "small_b7_bc96_d4": (replace(headport.PRESETS["b7_ultra"], roi_size=(32, 16), mask_size=(64, 32)), (64, 64)),
    "small_b0_ln_silu_noatt": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24),
                                       normalization_type="layernorm2d", activation_function="silu",
                                       use_attention_module=False), (96, 128)),
    "small_b0_bn_swish_beta": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24),
                                       activation_function="swish", activation_beta=1.5), (96, 128)),
    "small_b0_resize": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(40, 28)), (96, 128)),
}

# a13: no refinement flag -> PretrainedUNetGuidedSegmentationHead (rgb.py:715-727): the factory-default normalisation
# (layernorm2d, no attention) and a batchnorm + attention variant
GUIDED_CASES = {
    "small_b0_guided_ln": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), normalization_type="layernorm2d",
                                   use_attention_module=False, use_contour_detection=False, use_distance_transform=False), (96, 128)),
    "small_b0_guided_bn_att": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(40, 28), use_attention_module=True,
                                       use_contour_detection=False, use_distance_transform=False), (96, 128)),
}
# a13: use_pretrained_unet=False -> HierarchicalRGBSegmentationModel (rgb.py:298-439) with the V2 head (LayerNorm2d) or the
# refined head (batchnorm, attention, contour + distance)
STANDARD_CASES = {
    "small_std_v2_ln": (headport.PathConfig(roi_size=(16, 12), mask_size=(32, 24), normalization_type="layernorm2d",
                                            use_attention_module=False, use_contour_detection=False, use_distance_transform=False,
                                            use_pretrained_unet=False), (96, 128)),
    "small_std_refined_bn_att": (headport.PathConfig(roi_size=(16, 12), mask_size=(40, 28), normalization_type="batchnorm",
                                                     use_attention_module=True, use_pretrained_unet=False), (96, 128)),
}
# a7 flags that no headline preset enables: boundary refiner and sub-pixel re-decode (..._refinement.py:58-149, 218-252, 734-770)
REFINE_CASES = {
    "small_b0_boundary": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), use_boundary_refinement=True), (96, 128)),
    "small_b0_subpixel_boundary_ln": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(40, 28), use_boundary_refinement=True,
                                              use_subpixel_conv=True, use_contour_detection=False, use_distance_transform=False,
                                              normalization_type="layernorm2d", use_attention_module=False), (96, 128)),
    # a10: the other kinds of get_normalization_layer (normalization_comparison.py:159-206) through the full model
    "small_b0_groupnorm": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), normalization_type="groupnorm",
                                   normalization_groups=4, use_attention_module=True), (96, 128)),
    "small_b0_spatial_group": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), normalization_type="spatial_group",
                                       use_distance_transform=False, use_attention_module=False), (96, 128)),
    # instance / adaptive_instance (normalization_comparison.py:12-57,183,195): strict precision mode only on the B200 path
    "small_b0_instance": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), normalization_type="instance",
                                  use_attention_module=True, use_distance_transform=False), (96, 128)),
    "small_b0_adaptive_instance": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), normalization_type="adaptive_instance",
                                           use_attention_module=False, use_contour_detection=False), (96, 128)),
    # foreground_aware (:84-132, :200): instance norm blended per pixel by a learned detector; strict precision mode only
    "small_b0_foreground_aware": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), normalization_type="foreground_aware",
                                          use_attention_module=False, use_distance_transform=False), (96, 128)),
    "small_b0_mixed": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), normalization_type="mixed",
                               use_progressive_upsampling=True), (96, 128)),
    "small_b0_progressive": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(32, 24), use_progressive_upsampling=True,
                                     use_contour_detection=False, use_distance_transform=False), (96, 128)),
    "small_b0_progressive_ln_silu": (replace(headport.PRESETS["b0"], roi_size=(16, 12), mask_size=(40, 28), use_progressive_upsampling=True,
                                             use_boundary_refinement=True, normalization_type="layernorm2d", activation_function="silu",
                                             use_attention_module=False), (96, 128)),
}

# a1 multi_scale=True -> MultiScaleRGBSegmentationModel (rgb.py:777-922): default fusion/normalisation, and batchnorm + SiLU +
# adaptive fusion + attention with two scales
MULTISCALE_CASES = {
    "small_ms_concat_ln": (headport.PathConfig(mask_size=(56, 56), multi_scale=True, normalization_type="layernorm2d", use_attention_module=False,
                                               use_contour_detection=False, use_distance_transform=False, use_pretrained_unet=False), (96, 128)),
    "small_ms_adaptive_bn_silu_att": (headport.PathConfig(mask_size=(56, 56), multi_scale=True, roi_sizes=(("a", 28), ("b", 36)),
                                                          fusion_method="adaptive", normalization_type="batchnorm", activation_function="silu",
                                                          use_attention_module=True, use_contour_detection=False, use_distance_transform=False,
                                                          use_pretrained_unet=False), (96, 128)),
}
SMALL_CASES_ALL = {**SMALL_CASES, **GUIDED_CASES, **STANDARD_CASES, **MULTISCALE_CASES, **REFINE_CASES}

_FULL_KEYS = ("bg_fg_logits", "bg_fg_logits_low", "target_nontarget_logits", "contours", "distance_mask",
              "distance_map", "roi_features", "roi_patches")


def _np(t):
    return t.detach().cpu().numpy()


def _run_reference(cfg: headport.PathConfig, images, rois, seed=0, mode="stress"):
    model = refload.build_reference_model(**cfg.factory_kwargs())
    sd = paramfill.fill_state_dict(model.state_dict(), seed=seed, mode=mode)
    model.load_state_dict(sd)
    aligners = (model.roi_align_mask, model.roi_align_rgb) if hasattr(model, "roi_align_mask") else \
        tuple(model.roi_aligns.values()) if hasattr(model, "roi_aligns") else (model.roi_align,)
    for ra in aligners:      # export_onnx_advanced.py:80-98 mutates these
        ra.spatial_scale = cfg.spatial_scale
        ra.spatial_scale_h, ra.spatial_scale_w = cfg.spatial_scale
    with torch.no_grad():
        logits, aux = model(images, rois)
    return model, sd, logits, aux


def make_model_goldens():
    keys_json = {}
    for name, (cfg, (h, w)) in SMALL_CASES.items():
        images = synth_images(11, 2, h, w)
        rois = torch.cat([synth_rois(11, 2, 2), edge_rois(2)], 0)
        model, sd, logits, aux = _run_reference(cfg, images, rois)
        out = {"logits": _np(logits), "full_image_logits_ch0": _np(aux["full_image_logits"][:, 0]),
               "shared_features_sub": _np(aux["shared_features"][:, ::8]),
               "fg_attention_sub": _np(aux["fg_attention"][:, ::8])}
        for k in _FULL_KEYS:
            if k in aux:
                out[k] = _np(aux[k])
        Dil = refload.ref_class_from_script("export_hierarchical_instance_peopleseg_onnx.py", "MaskDilationModule")
        out["dilated1"] = _np(Dil(1)(logits))
        out["dilated2"] = _np(Dil(2)(logits))
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
        if "ln_" in name:
            keys_json[name] = {k: list(v.shape) for k, v in sd.items()}
        print(name, "logits", tuple(logits.shape), "absmax", float(logits.abs().max()),
              "argmax", torch.bincount(logits.argmax(1).flatten(), minlength=3).tolist())

    # state-dict key/shape contract of the three presets (+ the small cases above)
    for pname, cfg in headport.PRESETS.items():
        model = refload.build_reference_model(**cfg.factory_kwargs())
        keys_json["preset_" + pname] = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(GOLDEN, "state_dict_keys.json"), "w") as f:
        json.dump(keys_json, f)

    # BASELINE config 1: B0 std, 2x480x640, 8 ROIs (seeds: weights 0 / data 1)
    cfg = headport.PRESETS["b0"]
    images = synth_images(1, 2, 480, 640)
    rois = synth_rois(1, 2, 4)
    model, sd, logits, aux = _run_reference(cfg, images, rois)
    out = {"logits": _np(logits), "full_image_logits_ch0_s2": _np(aux["full_image_logits"][:, 0, ::2, ::2]),
           "shared_features_sub": _np(aux["shared_features"][:, ::8, ::4, ::4]),
           "fg_attention_sub": _np(aux["fg_attention"][:, ::8, ::4, ::4])}
    for k in ("bg_fg_logits_low", "target_nontarget_logits", "contours", "distance_map", "roi_features"):
        out[k] = _np(aux[k])
    np.savez_compressed(os.path.join(GOLDEN, "cfg1_b0.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
    print("cfg1_b0 logits absmax", float(logits.abs().max()),
          "argmax", torch.bincount(logits.argmax(1).flatten(), minlength=3).tolist())


def make_guided_goldens(only_case=None):
    """The guided-head variant; adds its key/shape tables to state_dict_keys.json without touching the other entries.
    only_case: regenerate that one case (``--only case:NAME``)."""
    path = os.path.join(GOLDEN, "state_dict_keys.json")
    keys_json = json.load(open(path))
    for name, (cfg, (h, w)) in {**GUIDED_CASES, **STANDARD_CASES, **MULTISCALE_CASES, **REFINE_CASES}.items():
        if only_case is not None and name != only_case:
            continue
        images = synth_images(11, 2, h, w)
        rois = torch.cat([synth_rois(11, 2, 2), edge_rois(2)], 0)
        model, sd, logits, aux = _run_reference(cfg, images, rois)
        out = {"logits": _np(logits)}
        if "full_image_logits" in aux:
            out["full_image_logits_ch0"] = _np(aux["full_image_logits"][:, 0])
        if "fg_attention" in aux:
            out["fg_attention_sub"] = _np(aux["fg_attention"][:, ::8])
        for k in ("bg_fg_logits", "target_nontarget_logits", "fg_prob", "pretrained_bg_fg_mask", "attention", "roi_features", "roi_patches",
                  "bg_fg_logits_low", "contours", "distance_mask", "distance_map"):
            if aux.get(k) is not None:
                out[k] = _np(aux[k])
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
        keys_json[name] = {k: list(v.shape) for k, v in sd.items()}
        print(name, "logits", tuple(logits.shape), "absmax", float(logits.abs().max()),
              "argmax", torch.bincount(logits.argmax(1).flatten(), minlength=3).tolist(), "aux", sorted(aux))
    with open(path, "w") as f:
        json.dump(keys_json, f)


def make_default_init_golden():
    """BASELINE config 1 again, with PyTorch-default-initialisation statistics ("random-init weights")."""
    cfg = headport.PRESETS["b0"]
    images = synth_images(1, 2, 480, 640)
    rois = synth_rois(1, 2, 4)
    model, sd, logits, aux = _run_reference(cfg, images, rois, mode="torch_default")
    out = {"logits": _np(logits), "bg_fg_logits_low": _np(aux["bg_fg_logits_low"]),
           "full_image_logits_ch0_s4": _np(aux["full_image_logits"][:, 0, ::4, ::4])}
    np.savez_compressed(os.path.join(GOLDEN, "cfg1_b0_default_init.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
    print("cfg1_b0_default_init logits range", float(logits.min()), float(logits.max()), "std", logits.std((0, 2, 3)).tolist(),
          "argmax", torch.bincount(logits.argmax(1).flatten(), minlength=3).tolist())


REAL_CASES = {
    # BASELINE configs[1] / [2] at their true ROI geometry ("no golden at 80x60->160x120 or 128x96->256x192"):
    # one 480x640 image, 4 ROIs, stress weights (seed 0), data seed 2 / 3 like SURVEY 8d
    "real_b1_enhanced": ("b1_enhanced", 2),
    "real_b7_ultra": ("b7_ultra", 3),
}


def real_case_inputs(name):
    preset, seed = REAL_CASES[name]
    return headport.PRESETS[preset], synth_images(seed, 1, 480, 640), synth_rois(seed, 1, 4)


def make_real_goldens():
    """B1 enhanced / B7 ultra presets end to end through the real reference modules at full ROI / mask resolution.  The stored
    tensors are sub-sampled where they are large (the file stays a few MB); the logits are stored whole."""
    for name in REAL_CASES:
        cfg, images, rois = real_case_inputs(name)
        model, sd, logits, aux = _run_reference(cfg, images, rois)
        out = {"logits": _np(logits), "full_image_logits_ch0_s4": _np(aux["full_image_logits"][:, 0, ::4, ::4]),
               "shared_features_sub": _np(aux["shared_features"][:, ::16, ::4, ::4]),
               "fg_attention_sub": _np(aux["fg_attention"][:, ::16, ::4, ::4]),
               "bg_fg_logits_low": _np(aux["bg_fg_logits_low"]), "roi_features": _np(aux["roi_features"]),
               "target_nontarget_logits_s2": _np(aux["target_nontarget_logits"][:, :, ::2, ::2]),
               "contours_s2": _np(aux["contours"][:, :, ::2, ::2]), "distance_map_s2": _np(aux["distance_map"][:, :, ::2, ::2])}
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
        print(name, "logits", tuple(logits.shape), "absmax", float(logits.abs().max()),
              "argmax", torch.bincount(logits.argmax(1).flatten(), minlength=3).tolist())


def make_roi_goldens():
    """DynamicRoIAlign itself (hed/dynamic_roi_align.py) on random feature maps, all conventions."""
    mod = refload.ref_import("dynamic_roi_align")
    g = torch.Generator().manual_seed(21)
    feat = torch.randn(3, 5, 37, 53, generator=g)
    rois = torch.cat([synth_rois(21, 3, 3), edge_rois(3)], 0)
    out = {"feat": _np(feat), "rois": _np(rois)}
    for tag, scale, aligned, (oh, ow) in [("a640", 640.0, True, (16, 12)), ("ahw", (37.0, 53.0), True, (16, 12)),
                                          ("u_hw", (37.0, 53.0), False, (7, 9)), ("a64", 64.0, True, (5, 3))]:
        ra = mod.DynamicRoIAlign(spatial_scale=scale, sampling_ratio=2, aligned=aligned)
        out[tag] = _np(ra(feat, rois, oh, ow))
    np.savez_compressed(os.path.join(GOLDEN, "roi_align.npz"), **out)
    print("roi_align goldens:", {k: v.shape for k, v in out.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if args.only in ("all", "roi"):
        make_roi_goldens()
    if args.only in ("all", "model"):
        make_model_goldens()
    if args.only in ("all", "model", "default"):
        make_default_init_golden()
    if args.only in ("all", "guided"):
        make_guided_goldens()
    if args.only.startswith("case:"):
        make_guided_goldens(args.only[5:])
    if args.only in ("all", "real"):
        make_real_goldens()
    if args.only in ("all", "post"):
        from . import make_golden_post
        make_golden_post.make_post_goldens(GOLDEN)
    if args.only in ("all", "post", "eval"):
        from . import make_golden_post
        make_golden_post.make_eval_goldens(GOLDEN)
    if args.only in ("all", "post", "post_variants"):
        from . import make_golden_post
        make_golden_post.make_post_variant_goldens(GOLDEN)


if __name__ == "__main__":
    main()
