"""Device-side metric core of ``train_utils.evaluate_model`` (hed/train_utils.py:109-404).

The reference argmaxes on the GPU, copies predictions and masks to the CPU and walks every sample with Python double loops
(:262-292).  Everything it reports from the predictions -- ``conf_matrix_total`` / ``_bg_target`` / ``_target_nontarget``,
``iou_class_k``, ``target_iou`` / ``miou``, ``detection_rate_*``, ``overall_accuracy``, ``target_precision/recall/f1``,
``instance_separation_accuracy`` -- is a function of nine integers per ROI, which one kernel produces
(``his_eval_confusion``); the remaining arithmetic is O(N) on the host in the reference's own order and types.
Loss terms are training-side and stay with the reference's loss modules.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import lib as _lib


@torch.no_grad()


@_lib.on_tensor_device
def roi_confusion_counts(logits: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] logits, [N,H,W] labels in {0,1,2} (uint8 or int64) -> int32 [N,3,3] counts[n, gt, pred]."""
    if not logits.is_cuda:
        raise _lib.HisError("roi_confusion_counts: CUDA tensors required (no CPU fallback)")
    x = logits.contiguous() if logits.dtype == torch.float32 else logits.float().contiguous()
    n, c, h, w = x.shape
    if c != 3 or masks.shape != (n, h, w):
        raise ValueError("expected logits [N,3,H,W] and masks [N,H,W]")
    m = masks.to(x.device)
    if m.dtype not in (torch.uint8, torch.int64):
        m = m.to(torch.int64)
    m = m.contiguous()
    out = torch.empty((n, 3, 3), dtype=torch.int32, device=x.device)
    L = _lib.load()
    st = torch.cuda.current_stream(x.device).cuda_stream
    hw = h * w
    for s0 in range(0, n, 65535):
        cnt = min(65535, n - s0)
        _lib.check(L.his_eval_confusion(x.data_ptr() + s0 * 3 * hw * 4, m.data_ptr() + s0 * hw * m.element_size(), 1 if m.dtype == torch.int64 else 0,
                                        cnt, h, w, out.data_ptr() + s0 * 9 * 4, st), "his_eval_confusion")
    return out


class EvalAccumulator:
    """Accumulates batches like the validation loop does and returns the reference's prediction metrics."""

    def __init__(self):
        self.conf_total = np.zeros((3, 3), np.int64)
        self.conf_bg_target = np.zeros((2, 2), np.int64)
        self.conf_tn = np.zeros((2, 2), np.int64)
        self.class_ious: Dict[int, List[float]] = {0: [], 1: [], 2: []}

    def update(self, logits: torch.Tensor, masks: torch.Tensor):
        c = roi_confusion_counts(logits, masks).cpu().numpy().astype(np.int64)          # [N,3,3], one small D2H copy per batch
        tot = c.sum(0)
        self.conf_total += tot
        # background-vs-target (train_utils.py:262-271): class 1 against the rest
        self.conf_bg_target += np.array([[tot.sum() - tot[1].sum() - tot[:, 1].sum() + tot[1, 1], tot[0, 1] + tot[2, 1]],
                                         [tot[1, 0] + tot[1, 2], tot[1, 1]]], np.int64)
        # target-vs-non-target on ground-truth foreground pixels (:273-280): "is class 2" for label and prediction
        self.conf_tn += np.array([[tot[1, 0] + tot[1, 1], tot[1, 2]], [tot[2, 0] + tot[2, 1], tot[2, 2]]], np.int64)
        # calculate_iou (:14-22) per class and sample, in float32 like the reference's tensors
        for k in range(3):
            inter = c[:, k, k].astype(np.float32)
            union = (c[:, k, :].sum(1) + c[:, :, k].sum(1) - c[:, k, k]).astype(np.float32)
            iou = np.where(union == 0, np.float32(1.0), inter / np.maximum(union, np.float32(1.0)))
            self.class_ious[k] += [float(v) for v in iou]

    def compute(self) -> dict:
        m: dict = {}
        for k in range(3):
            v = self.class_ious[k]
            m[f"iou_class_{k}"] = sum(v) / len(v) if v else 0.0
        t = self.class_ious[1]
        m["target_iou"] = sum(t) / len(t) if t else 0.0
        m["miou"] = m["target_iou"]
        for thr in (0.5, 0.7):                                                   # calculate_detection_metrics (:85-106)
            m[f"detection_rate_{thr}"] = float((np.array(t) > thr).mean()) if t else 0.0
        ct, cb, cn = self.conf_total, self.conf_bg_target, self.conf_tn
        if ct.sum() > 0:
            m["overall_accuracy"] = np.diag(ct).sum() / ct.sum()
        if cb.sum() > 0:
            tp, fp, fn = cb[1, 1], cb[0, 1], cb[1, 0]
            m["target_precision"] = tp / (tp + fp) if tp + fp > 0 else 0.0
            m["target_recall"] = tp / (tp + fn) if tp + fn > 0 else 0.0
            pr = m["target_precision"] + m["target_recall"]
            m["target_f1"] = 2 * (m["target_precision"] * m["target_recall"]) / pr if pr > 0 else 0.0
        if cn.sum() > 0:
            m["instance_separation_accuracy"] = np.diag(cn).sum() / cn.sum()
        m["conf_matrix_total"], m["conf_matrix_bg_target"], m["conf_matrix_target_nontarget"] = ct.copy(), cb.copy(), cn.copy()
        return m


def evaluate_predictions(logits: torch.Tensor, masks: torch.Tensor) -> dict:
    acc = EvalAccumulator()
    acc.update(logits, masks)
    return acc.compute()
