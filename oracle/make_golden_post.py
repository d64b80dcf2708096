"""Oracle (test infrastructure): golden vectors of the post-processing family from the REAL reference modules/functions."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

from human_instance_segmentation_b200.synthetic import synth_rois
from . import refload


def blob_masks(seed: int, n: int, h: int, w: int) -> torch.Tensor:
    """SURVEY §8d cfg 5: (rand > 0.5) smoothed by a 15x15 box blur, re-thresholded -> blob-like binary masks."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, 1, h, w, generator=g) > 0.5).float()
    x = F.avg_pool2d(F.pad(x, (7, 7, 7, 7), mode="replicate"), 15, 1)
    return (x > 0.5).float()


def all_3x3_patterns() -> torch.Tensor:
    """Every binary 3x3 neighbourhood once, centre pixels on a 5-pixel pitch (the tie-sensitive cases of SURVEY §8 a16)."""
    img = torch.zeros(1, 1, 16 * 5 + 2, 32 * 5 + 2)
    for p in range(512):
        r, c = divmod(p, 32)
        for b in range(9):
            if (p >> b) & 1:
                img[0, 0, 1 + r * 5 + b // 3, 1 + c * 5 + b % 3] = 1.0
    return img


def make_post_variant_goldens(golden_dir: str):
    """a16/a17 variants: outputs of the reference's own classes (export_edge_smoothing_onnx.py, hed/edge_smoothing.py,
    hed/bilateral_filter.py) on seeded inputs -> tests/golden/post_variants.npz."""
    es = refload.ref_import("edge_smoothing")
    bf = refload.ref_import("bilateral_filter")
    cls = {n: refload.ref_class_from_script("export_edge_smoothing_onnx.py", n)
           for n in ("DirectionalEdgeSmoothing", "AdaptiveEdgeSmoothing", "OptimizedEdgeSmoothing")}
    out = {}
    masks = blob_masks(7, 3, 96, 128)
    g = torch.Generator().manual_seed(8)
    gray = torch.rand(2, 2, 40, 56, generator=g)                       # float images for the value filters
    guide = (gray + 0.1 * torch.randn(gray.shape, generator=g)).contiguous()
    logits = torch.randn(2, 3, 48, 64, generator=g) * 2
    logits = F.avg_pool2d(F.pad(logits, (3, 3, 3, 3), mode="replicate"), 7, 1).contiguous()      # blob-like argmax regions
    probs5 = torch.rand(2, 5, 48, 64, generator=g)
    bs, sens, thr = torch.tensor([[1.0], [3.0], [5.0]]), torch.tensor([[0.5], [1.0], [2.0]]), torch.tensor([[0.3], [0.5], [0.7]])
    out.update(masks=masks.numpy(), gray=gray.numpy(), guide=guide.numpy(), logits=logits.numpy(), probs5=probs5.numpy(),
               ad_bs=bs.numpy(), ad_sens=sens.numpy(), ad_thr=thr.numpy(), patterns=all_3x3_patterns().numpy())
    with torch.no_grad():
        out["directional"] = cls["DirectionalEdgeSmoothing"]()(masks).numpy()
        out["adaptive"] = cls["AdaptiveEdgeSmoothing"]()(masks, bs, sens, thr).numpy()
        out["optimized_fp32"] = cls["OptimizedEdgeSmoothing"](use_fp16=False)(masks).numpy()
        out["optimized_fp32_patterns"] = cls["OptimizedEdgeSmoothing"](use_fp16=False)(all_3x3_patterns()).numpy()
        # use_fp16=True needs half weights (the exporter converts the module: export_edge_smoothing_onnx.py model.half())
        out["optimized_fp16"] = cls["OptimizedEdgeSmoothing"](use_fp16=True).half()(masks).float().numpy()
        mc = es.MultiClassEdgeSmoothing(device="cpu")
        out["multiclass3"] = mc.smooth_predictions(logits).numpy()
        out["multiclass3_softmax_it2"] = es.MultiClassEdgeSmoothing(0.5, 3.0, 2, device="cpu").smooth_predictions(logits, apply_softmax=True).numpy()
        out["multiclass5"] = mc.smooth_predictions(probs5).numpy()
        out["bilateral_exact"] = bf.BilateralFilter()(gray[:, :, :12, :16].contiguous()).numpy()       # Python triple loop: tiny crop
        out["bilateral_exact_k3"] = bf.BilateralFilter(3, 0.8, 0.3)(gray[:1, :, :10, :12].contiguous()).numpy()
        out["bilateral_fast"] = bf.FastBilateralFilter()(gray).numpy()
        out["bilateral_fast_k7_it3"] = bf.FastBilateralFilter(7, 1.5, 0.2, 3)(gray).numpy()
        out["edge_preserving"] = bf.EdgePreservingFilter()(gray).numpy()
        out["edge_preserving_guided_r3"] = bf.EdgePreservingFilter(3, 0.05)(gray, guide).numpy()
    np.savez_compressed(os.path.join(golden_dir, "post_variants.npz"), **out)
    print("post variant goldens:", {k: v.shape for k, v in out.items()})


def make_post_goldens(golden_dir: str):
    es = refload.ref_import("edge_smoothing")
    bf = refload.ref_import("bilateral_filter")
    out = {}
    masks = blob_masks(5, 3, 96, 128)
    noisy = (masks + 0.15 * torch.randn(masks.shape, generator=torch.Generator().manual_seed(6))).contiguous()   # float-valued, outside [0,1]
    out["masks"], out["noisy"] = masks.numpy(), noisy.numpy()
    with torch.no_grad():
        out["edge_smooth"] = es.BinaryMaskEdgeSmoothing()(masks).numpy()
        out["edge_smooth_t04_s2"] = es.BinaryMaskEdgeSmoothing(0.4, 2.0)(masks).numpy()
        pat = all_3x3_patterns()
        out["patterns"] = pat.numpy()
        out["edge_smooth_patterns"] = es.BinaryMaskEdgeSmoothing()(pat).numpy()
        out["binary_bilateral"] = bf.BinaryMaskBilateralFilter()(masks).numpy()
        out["binary_bilateral_noisy"] = bf.BinaryMaskBilateralFilter()(noisy).numpy()
        out["binary_bilateral_k5_it3"] = bf.BinaryMaskBilateralFilter(5, 1.0, 0.5, 3)(masks).numpy()
        out["morph_bilateral"] = bf.MorphologicalBilateralFilter()(masks).numpy()
        out["morph_bilateral_noisy"] = bf.MorphologicalBilateralFilter()(noisy).numpy()
    # paste-back through the reference script's own functions
    fn = refload.ref_functions_from_script("test_hierarchical_instance_peopleseg_onnx.py", ["denormalize_bbox", "process_mask_output"])
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(7, 3, 32, 24, generator=g) * 2
    rois = synth_rois(9, 1, 7)
    rois[3, 1:] = torch.tensor([0.7, 0.699999988, 0.95, 1.0])     # products that land next to an integer
    H, W = 480, 640
    res = fn["process_mask_output"](logits.numpy(), rois.numpy(), W, H, 0.5)
    canvas = np.zeros((1, H, W), np.int32)
    target = np.zeros((7, 32, 24), np.uint8)
    for i, r in enumerate(res):
        x1, y1, x2, y2 = r["bbox"]
        full = np.zeros((H, W), np.uint8)
        if x2 > x1 and y2 > y1:
            full[y1:y2, x1:x2] = r["mask"]
        canvas[0][full > 0] = i + 1
        sm = F.softmax(logits[i], 0).numpy()
        target[i] = ((r["raw_mask"] == 1) & (sm.max(0) > 0.5)).astype(np.uint8)
    assert len(res) == 7
    out["paste_logits"], out["paste_rois"], out["paste_canvas"], out["paste_target"] = logits.numpy(), rois.numpy(), canvas, target
    np.savez_compressed(os.path.join(golden_dir, "post.npz"), **out)
    print("post goldens:", {k: v.shape for k, v in out.items()})


def make_eval_goldens(golden_dir: str):
    """evaluate_model's prediction metrics with the REFERENCE's helper functions (hed/train_utils.py:14-47,85-106) driving
    the restated loop of oracle/evalport.py -> tests/golden/eval_metrics.npz."""
    from . import evalport
    fns = refload.ref_functions_from_script("src/human_edge_detection/train_utils.py",
                                            ["calculate_iou", "calculate_confusion_matrix", "calculate_detection_metrics"])
    m = evalport.evaluate(evalport.synth_eval_batches(), helpers=fns)
    np.savez_compressed(os.path.join(golden_dir, "eval_metrics.npz"), **{k: np.asarray(v) for k, v in m.items()})
    print("eval goldens:", {k: (v if np.ndim(v) == 0 else np.asarray(v).tolist()) for k, v in m.items()})
