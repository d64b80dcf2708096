"""the prediction metrics of ``train_utils.evaluate_model`` (hed/train_utils.py:109-404).
CPU: the oracle port against the golden produced with the reference's own helper functions.  GPU: the device-side
confusion kernel + host summary against the same golden (integers bit-exact, IoUs the same float32 quotients)."""
import numpy as np
import pytest
import torch

from oracle import evalport
from tests import common


def _same(m, g):
    for k in ("conf_matrix_total", "conf_matrix_bg_target", "conf_matrix_target_nontarget"):
        assert np.array_equal(np.asarray(m[k]), g[k].numpy()), k
    for k in ("iou_class_0", "iou_class_1", "iou_class_2", "target_iou", "miou", "detection_rate_0.5", "detection_rate_0.7",
              "overall_accuracy", "target_precision", "target_recall", "target_f1", "instance_separation_accuracy"):
        assert abs(float(m[k]) - float(g[k])) <= 1e-12, k


def test_eval_port_matches_reference_helpers_golden():
    _same(evalport.evaluate(evalport.synth_eval_batches()), common.golden("eval_metrics"))


@pytest.mark.gpu
def test_device_side_eval_metrics_match_reference():
    from human_instance_segmentation_b200 import metrics
    acc = metrics.EvalAccumulator()
    for logits, masks in evalport.synth_eval_batches():
        acc.update(logits.cuda(), masks.cuda())                     # int64 labels, as the dataloader yields them
    _same(acc.compute(), common.golden("eval_metrics"))
    # uint8 labels, one call, against the oracle on fresh data at the B0 mask size; N = 0
    batches = evalport.synth_eval_batches(seed=31, n_batches=1, n=9, h=128, w=96)
    got = metrics.evaluate_predictions(batches[0][0].cuda(), batches[0][1].to(torch.uint8).cuda())
    want = evalport.evaluate(batches)
    assert np.array_equal(got["conf_matrix_total"], want["conf_matrix_total"])
    assert abs(got["target_iou"] - want["target_iou"]) <= 1e-12 and abs(got["iou_class_2"] - want["iou_class_2"]) <= 1e-12
    assert metrics.roi_confusion_counts(batches[0][0][:0].cuda(), batches[0][1][:0].cuda()).shape == (0, 3, 3)
