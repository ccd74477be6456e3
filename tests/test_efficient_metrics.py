"""Streaming count metrics: docstring oracles of the reference (efficient_metrics_base.py:464-471, 733-741) plus
hand-computed Dice values; EMA / transform wrappers."""

import pytest
import torch

from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.compound_metrics import EmaMetric, TransformsMetric
from fl4health_b200.metrics.efficient_metrics import BinaryDice, MultiClassDice
from fl4health_b200.metrics.efficient_metrics_base import (
    BinaryClassificationMetric,
    ClassificationOutcome,
    MultiClassificationMetric,
)
from fl4health_b200.metrics.metrics_utils import compute_dice_on_count_tensors, threshold_tensor
from fl4health_b200.metrics.utils import align_pred_and_target_shapes, infer_label_dim


class _BinaryCounts(BinaryClassificationMetric):
    def compute_from_counts(self, true_positives, false_positives, true_negatives, false_negatives):
        return {"recall": (true_positives.sum() / (true_positives.sum() + false_negatives.sum())).item()}


class _MultiCounts(MultiClassificationMetric):
    def compute_from_counts(self, true_positives, false_positives, true_negatives, false_negatives):
        return {"tp": true_positives.sum().item()}


def test_binary_counts_batch_and_label_dims() -> None:
    metric = _BinaryCounts("m", label_dim=0, batch_dim=1)
    p = torch.tensor([[[0, 0, 0, 1], [1, 1, 1, 1]]])
    t = torch.tensor([[[0, 0, 1, 0], [1, 1, 1, 1]]])
    tp, fp, tn, fn = metric.count_tp_fp_tn_fn(p, t)
    assert tp.tolist() == [[0.0], [4.0]] and tn.tolist() == [[2.0], [0.0]]
    assert fp.tolist() == [[1.0], [0.0]] and fn.tolist() == [[1.0], [0.0]]
    metric.update(p, t)
    metric.update(p, t)
    assert metric.true_positives.shape == (4, 1)
    assert metric.compute("val")["val - recall"] == pytest.approx(8 / 10)
    flipped = _BinaryCounts("m", label_dim=0, batch_dim=1, pos_label=0)
    tp0, fp0, tn0, fn0 = flipped.count_tp_fp_tn_fn(p, t)
    assert torch.equal(tp0, tn) and torch.equal(fn0, fp)


def test_multi_counts_batch_first() -> None:
    metric = _MultiCounts("m", label_dim=0, batch_dim=1)
    p = torch.tensor([[[1.0, 1.0, 1.0, 0.0]], [[0.0, 0.0, 0.0, 1.0]]])
    t = torch.tensor([[[1.0, 1.0, 0.0, 0.0]], [[0.0, 0.0, 1.0, 1.0]]])
    tp, fp, tn, fn = metric.count_tp_fp_tn_fn(p, t)
    assert tp.tolist() == [[2.0, 1.0]] and tn.tolist() == [[1.0, 2.0]]
    assert fp.tolist() == [[1.0, 0.0]] and fn.tolist() == [[0.0, 1.0]]
    discard = _MultiCounts("m", label_dim=0, discard={ClassificationOutcome.TRUE_NEGATIVE})
    assert discard.count_tp_fp_tn_fn(p, t)[2].numel() == 0


def test_soft_counts_and_thresholds() -> None:
    metric = _BinaryCounts("m")
    tp, fp, tn, fn = metric.count_tp_fp_tn_fn(torch.tensor([0.8, 0.3]), torch.tensor([1.0, 0.0]))
    assert tp.item() == pytest.approx(0.8) and fn.item() == pytest.approx(0.2)
    assert fp.item() == pytest.approx(0.3) and tn.item() == pytest.approx(0.7)
    assert threshold_tensor(torch.tensor([0.2, 0.7]), 0.5).tolist() == [0.0, 1.0]
    assert threshold_tensor(torch.tensor([[0.2, 0.7, 0.1], [0.5, 0.1, 0.4]]), 1).tolist() == [[0, 1, 0], [1, 0, 0]]
    with pytest.raises(AssertionError):
        metric.count_tp_fp_tn_fn(torch.tensor([1.5]), torch.tensor([1.0]))


def test_shape_alignment() -> None:
    preds = torch.tensor([[0.1, 0.2, 0.7], [0.9, 0.1, 0.0]])
    p, t = align_pred_and_target_shapes(preds, torch.tensor([[2], [1]]))
    assert t.tolist() == [[0, 0, 1], [0, 1, 0]] and p is preds
    p, t = align_pred_and_target_shapes(preds, torch.tensor([2, 1]))
    assert t.tolist() == [[0, 0, 1], [0, 1, 0]]
    assert infer_label_dim(torch.zeros(4, 3, 8, 8), torch.zeros(4, 8, 8)) == 1
    with pytest.raises(AssertionError):
        infer_label_dim(torch.zeros(5, 5, 3), torch.zeros(5, 3))


def test_dice_values() -> None:
    assert compute_dice_on_count_tensors(torch.tensor([2.0, 0.0]), torch.tensor([1.0, 0.0]), torch.tensor([1.0, 0.0]), None).tolist() == [
        pytest.approx(4 / 6)]
    assert compute_dice_on_count_tensors(torch.tensor([2.0, 0.0]), torch.tensor([1.0, 0.0]), torch.tensor([1.0, 0.0]), 1.0).tolist() == [
        pytest.approx(4 / 6), 1.0]
    binary = BinaryDice(batch_dim=None, threshold=0.5)
    binary.update(torch.tensor([0.9, 0.8, 0.2, 0.1]), torch.tensor([1, 0, 1, 0]))
    assert binary.compute()["BinaryDice"] == pytest.approx(2 * 1 / (2 * 1 + 1 + 1))
    binary.update(torch.tensor([0.9, 0.8]), torch.tensor([1, 1]))
    assert binary.compute()["BinaryDice"] == pytest.approx(2 * 3 / (2 * 3 + 1 + 1))
    binary.clear()
    assert not binary.counts_initialized
    per_sample = BinaryDice(batch_dim=0, threshold=0.5, zero_division=None)
    per_sample.update(torch.tensor([[0.9, 0.8], [0.1, 0.2]]), torch.tensor([[1, 0], [0, 0]]))  # second sample: only TN -> dropped
    assert per_sample.compute()["BinaryDice"] == pytest.approx(2 / 3)
    neg = BinaryDice(batch_dim=None, threshold=0.5, pos_label=0)
    neg.update(torch.tensor([0.9, 0.8, 0.2, 0.1]), torch.tensor([1, 0, 1, 0]))
    assert neg.compute()["BinaryDice"] == pytest.approx(0.5)
    multi = MultiClassDice(batch_dim=None, label_dim=1, threshold=1)
    logits = torch.tensor([[0.7, 0.2, 0.1], [0.1, 0.8, 0.1], [0.3, 0.3, 0.4], [0.6, 0.3, 0.1]])
    multi.update(logits, torch.tensor([0, 1, 1, 0]))
    # class0: tp2 fp0 fn0 -> 1 ; class1: tp1 fp0 fn1 -> 2/3 ; class2: tp0 fp1 fn0 -> 0
    assert multi.compute("val")["val - MultiClassDice"] == pytest.approx((1 + 2 / 3 + 0) / 3)
    assert multi(logits, torch.tensor([0, 1, 1, 0])) == pytest.approx((1 + 2 / 3 + 0) / 3)
    no_bg = MultiClassDice(batch_dim=None, label_dim=1, threshold=1, ignore_background=1)
    no_bg.update(logits, torch.tensor([0, 1, 1, 0]))
    assert no_bg.compute()["MultiClassDice"] == pytest.approx((2 / 3 + 0) / 2)


def test_ema_and_transforms_metrics() -> None:
    ema = EmaMetric(Accuracy(), 0.1)
    ema.update(torch.tensor([[0.0, 1.0], [1.0, 0.0], [0.0, 1.0]]), torch.tensor([1, 1, 1]))
    first = ema.compute()["EMA_accuracy"]
    assert first == pytest.approx(2 / 3)
    ema.clear()
    ema.update(torch.tensor([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), torch.tensor([1, 1, 1]))
    second = ema.compute()["EMA_accuracy"]
    assert second == pytest.approx(0.9 * (2 / 3) + 0.1 * (1 / 3))
    wrapped = TransformsMetric(Accuracy(), pred_transforms=[lambda p: 1 - p], target_transforms=[lambda t: t.long()])
    wrapped.update(torch.tensor([[0.9, 0.1], [0.2, 0.8]]), torch.tensor([1.0, 0.0]))
    assert wrapped.compute("t")["t - accuracy"] == pytest.approx(1.0)
    wrapped.clear()
