"""Shared helpers for the parity tests: seeded inputs, procedural weights, golden loading."""
import json
import os

import numpy as np
import torch

from oracle import headport, paramfill
from oracle.make_golden import (GUIDED_CASES, MULTISCALE_CASES, REAL_CASES, REFINE_CASES, SMALL_CASES, SMALL_CASES_ALL, STANDARD_CASES,  # noqa: F401
                                edge_rois, real_case_inputs, synth_images, synth_rois)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}


def golden_keys():
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        return json.load(f)


def small_case_inputs(name):
    cfg, (h, w) = SMALL_CASES_ALL[name]
    images = synth_images(11, 2, h, w)
    rois = torch.cat([synth_rois(11, 2, 2), edge_rois(2)], 0)
    return cfg, images, rois


def cfg1_inputs():
    return headport.PRESETS["b0"], synth_images(1, 2, 480, 640), synth_rois(1, 2, 4)


def procedural_state(shapes: dict, seed=0, weights_path="ext_extractor/best_model_b0_0.8741.pth", mode="stress"):
    """Builds the procedural state dict for a key->shape table (tests/golden/state_dict_keys.json)."""
    sd = {}
    for k, shp in shapes.items():
        dt = torch.int64 if k.endswith("num_batches_tracked") else torch.float32
        sd[k] = torch.zeros(shp, dtype=dt)
    # values the reference pins in its constructors
    if "pretrained_unet.output_conv.weight" in sd:
        sd["pretrained_unet.output_conv.weight"][:, 0, 0, 0] = torch.tensor([1.0, -1.0])
    m, s = headport.imagenet_or_half_norm(weights_path)
    if "pretrained_unet.model.norm_mean" in sd:
        sd["pretrained_unet.model.norm_mean"] = torch.tensor(m).view(1, 3, 1, 1)
        sd["pretrained_unet.model.norm_std"] = torch.tensor(s).view(1, 3, 1, 1)
    for k in sd:
        if k.endswith("distance_decoder.threshold"):
            sd[k] = torch.tensor(0.3)
        if k.endswith("boundary_refiner.blend_weight"):
            sd[k] = torch.tensor(0.01)           # ..._refinement.py:90 (scalars keep their constructor value)
    return paramfill.fill_state_dict(sd, seed=seed, mode=mode)


def shapes_for_case(name):
    """Key/shape table of a small case: same architecture as a preset except LN variants."""
    keys = golden_keys()
    if name in keys:
        return keys[name]
    cfg = SMALL_CASES_ALL[name][0]
    for pname, p in headport.PRESETS.items():
        if (p.encoder_name, p.hierarchical_base_channels, p.hierarchical_depth) == \
                (cfg.encoder_name, cfg.hierarchical_base_channels, cfg.hierarchical_depth):
            return keys["preset_" + pname]
    raise KeyError(name)


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def argmax_agreement(a, b):
    return float((a.argmax(1) == b.argmax(1)).float().mean())
