"""Oracle (test infrastructure): the prediction-metric part of ``evaluate_model`` (hed/train_utils.py:109-404) restated for
one or more batches of (logits [N,3,H,W], masks [N,H,W]).  The helper functions are line-for-line restatements of
``calculate_iou`` (:14-22), ``calculate_confusion_matrix`` (:25-47) and ``calculate_detection_metrics`` (:85-106); the loop
follows :249-292 and the summary :325-375.  ``helpers`` lets the golden generator swap in the reference's own functions."""
from __future__ import annotations

import numpy as np
import torch


def calculate_iou(pred, target):
    inter = (pred & target).float().sum()
    union = (pred | target).float().sum()
    if union == 0:
        return 1.0 if inter == 0 else 0.0
    return (inter / union).item()


def calculate_confusion_matrix(pred, target, num_classes=3):
    pf, tf = pred.view(-1), target.view(-1)
    cm = torch.zeros(num_classes, num_classes, dtype=torch.int64)
    for t in range(num_classes):
        for p in range(num_classes):
            cm[t, p] = ((tf == t) & (pf == p)).sum()
    return cm


def calculate_detection_metrics(ious, thresholds=(0.5, 0.7)):
    if not ious:
        return {f"detection_rate_{t}": 0.0 for t in thresholds}
    arr = np.array(ious)
    return {f"detection_rate_{t}": (arr > t).mean() for t in thresholds}


def evaluate(batches, helpers=None) -> dict:
    h = helpers or {"calculate_iou": calculate_iou, "calculate_confusion_matrix": calculate_confusion_matrix,
                    "calculate_detection_metrics": calculate_detection_metrics}
    class_ious = {i: [] for i in range(3)}
    target_ious = []
    ct = torch.zeros(3, 3, dtype=torch.int64)
    cb = torch.zeros(2, 2, dtype=torch.int64)
    cn = torch.zeros(2, 2, dtype=torch.int64)
    for logits, masks in batches:
        pred = logits.argmax(dim=1)
        ct += h["calculate_confusion_matrix"](pred, masks)
        pb, tb = (pred == 1).long(), (masks == 1).long()
        for i in range(pred.shape[0]):
            for t in range(2):
                for p in range(2):
                    cb[t, p] += ((tb[i] == t) & (pb[i] == p)).sum().item()
            fg = masks[i] > 0
            if fg.any():
                ptn, ttn = (pred[i] == 2).long()[fg], (masks[i] == 2).long()[fg]
                for t in range(2):
                    for p in range(2):
                        cn[t, p] += ((ttn == t) & (ptn == p)).sum().item()
        for cls in range(3):
            pm, tm = pred == cls, masks == cls
            for i in range(pm.shape[0]):
                iou = h["calculate_iou"](pm[i], tm[i])
                class_ious[cls].append(iou)
                if cls == 1:
                    target_ious.append(iou)
    m = {}
    for cls in range(3):
        m[f"iou_class_{cls}"] = sum(class_ious[cls]) / len(class_ious[cls]) if class_ious[cls] else 0.0
    m["target_iou"] = sum(target_ious) / len(target_ious) if target_ious else 0.0
    m["miou"] = m["target_iou"]
    m.update(h["calculate_detection_metrics"](target_ious))
    ctn, cbn, cnn = ct.numpy(), cb.numpy(), cn.numpy()
    if ctn.sum() > 0:
        m["overall_accuracy"] = np.diag(ctn).sum() / ctn.sum()
    if cbn.sum() > 0:
        tp, fp, fn = cbn[1, 1], cbn[0, 1], cbn[1, 0]
        m["target_precision"] = tp / (tp + fp) if tp + fp > 0 else 0.0
        m["target_recall"] = tp / (tp + fn) if tp + fn > 0 else 0.0
        pr = m["target_precision"] + m["target_recall"]
        m["target_f1"] = 2 * (m["target_precision"] * m["target_recall"]) / pr if pr > 0 else 0.0
    if cnn.sum() > 0:
        m["instance_separation_accuracy"] = np.diag(cnn).sum() / cnn.sum()
    m["conf_matrix_total"], m["conf_matrix_bg_target"], m["conf_matrix_target_nontarget"] = ctn, cbn, cnn
    return m


def synth_eval_batches(seed=12, n_batches=2, n=5, h=32, w=24):
    """Seeded (logits, masks) with blob-like labels; one ROI with an empty class (union == 0 branch)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(n_batches):
        base = torch.randn(n, 3, h // 4, w // 4, generator=g)
        masks = torch.nn.functional.interpolate(base, size=(h, w), mode="bilinear").argmax(1)
        logits = torch.nn.functional.interpolate(base + 0.8 * torch.randn(n, 3, h // 4, w // 4, generator=g), size=(h, w), mode="bilinear")
        logits = logits.contiguous()
        if b == 0:
            masks[0] = 0                          # no foreground at all in this ROI
            logits[0, 0] += 10.0                  # ... and predicted as such: classes 1, 2 have union 0 -> IoU 1.0
            logits[1, :, :4] = 0.25               # exact ties: argmax must return the first maximum
        out.append((logits, masks))
    return out
