"""Loss containers / meters and the classification metrics against scikit-learn (mirrors the reference's
tests/utils/losses_test.py and tests/metrics/*): the streaming device-side implementations must give the numbers the
reference's accumulate-then-sklearn versions give."""

import numpy as np
import pytest
import torch
from sklearn import metrics as sk

from fl4health_b200.metrics import F1, Accuracy, BalancedAccuracy, BinarySoftDiceCoefficient, RocAuc
from fl4health_b200.metrics.metric_managers import MetricManager
from fl4health_b200.utils.losses import EvaluationLosses, LossMeter, LossMeterType, TrainingLosses


def test_loss_meter_average_and_accumulation() -> None:
    average = LossMeter(LossMeterType.AVERAGE, TrainingLosses)
    accumulate = LossMeter(LossMeterType.ACCUMULATION, TrainingLosses)
    for value in (1.0, 2.0, 6.0):
        losses = TrainingLosses(torch.tensor(value), {"extra": torch.tensor(2 * value)})
        average.update(losses)
        accumulate.update(losses)
    assert average.compute().as_dict() == {"extra": 6.0, "backward": 3.0}
    assert accumulate.compute().as_dict() == {"extra": 18.0, "backward": 9.0}
    average.clear()
    with pytest.raises(AssertionError, match="empty loss meter"):
        average.compute()
    # dict-valued backward losses (multi-optimizer clients) keep their keys
    multi = LossMeter(LossMeterType.AVERAGE, TrainingLosses)
    multi.update(TrainingLosses({"global": torch.tensor(1.0), "local": torch.tensor(3.0)}))
    multi.update(TrainingLosses({"global": torch.tensor(3.0), "local": torch.tensor(5.0)}))
    assert multi.compute().as_dict() == {"global": 2.0, "local": 4.0}
    evaluation = LossMeter(LossMeterType.AVERAGE, EvaluationLosses)
    evaluation.update(EvaluationLosses(torch.tensor(4.0)))
    evaluation.update(EvaluationLosses(torch.tensor(2.0)))
    assert evaluation.compute().as_dict() == {"checkpoint": 3.0}
    # graph-replay bookkeeping: accumulate() without counting, then mark_step
    replayed = LossMeter(LossMeterType.AVERAGE, EvaluationLosses)
    replayed.accumulate(EvaluationLosses(torch.tensor(8.0)))
    replayed.mark_step()
    assert replayed.compute().as_dict() == {"checkpoint": 8.0}


def _batches(n_batches: int, classes: int, seed: int):
    gen = torch.Generator().manual_seed(seed)
    logits = [torch.randn(17, classes, generator=gen) for _ in range(n_batches)]
    targets = [torch.randint(0, classes, (17,), generator=gen) for _ in range(n_batches)]
    return logits, targets


def test_streaming_classification_metrics_match_sklearn() -> None:
    logits, targets = _batches(5, 4, seed=3)
    y_true = torch.cat(targets).numpy()
    y_pred = torch.cat(logits).argmax(dim=1).numpy()
    for metric, expected in ((Accuracy(), sk.accuracy_score(y_true, y_pred)),
                             (BalancedAccuracy(), sk.balanced_accuracy_score(y_true, y_pred)),
                             (F1(average="weighted"), sk.f1_score(y_true, y_pred, average="weighted")),
                             (F1(average="macro"), sk.f1_score(y_true, y_pred, average="macro"))):
        for batch_logits, batch_targets in zip(logits, targets):
            metric.update(batch_logits, batch_targets)
        (value,) = metric.compute("val - prediction").values()
        assert abs(float(value) - expected) < 1e-6, type(metric).__name__
        metric.clear()
    # ROC-AUC, weighted one-vs-rest on the softmax of the logits (the metric applies the softmax itself)
    probs = torch.softmax(torch.cat(logits), dim=1)
    auc = RocAuc()
    for batch_logits, batch_targets in zip(logits, targets):
        auc.update(batch_logits, batch_targets)
    (value,) = auc.compute().values()
    assert abs(float(value) - sk.roc_auc_score(y_true, probs.numpy(), average="weighted", multi_class="ovr")) < 1e-5


def test_binary_soft_dice_and_metric_manager_keys() -> None:
    # [batch, channel, x, y, z] volumes; thresholded at 0.5 by default -> dice of the hard prediction
    pred = torch.zeros(2, 1, 2, 2, 2)
    target = torch.zeros(2, 1, 2, 2, 2)
    pred[0, 0, 0] = 0.9          # 4 voxels predicted, 2 of them true
    target[0, 0, 0, 0] = 1.0
    pred[1, 0, :, 0, 0] = 0.8    # 2 voxels predicted, both true, plus 2 missed
    target[1, 0, :, 0] = 1.0
    expected = np.mean([2 / (0.5 * (4 + 2)), 2 / (0.5 * (2 + 4))])
    assert abs(float(BinarySoftDiceCoefficient()(pred, target)) - expected) < 1e-6
    soft = BinarySoftDiceCoefficient(logits_threshold=None)(pred, target)
    assert 0.0 < float(soft) < expected  # un-thresholded scores shrink the intersection
    manager = MetricManager([Accuracy()], "val")
    logits, targets = _batches(2, 3, seed=9)
    manager.update({"prediction": logits[0], "aux": logits[1]}, targets[0])
    result = manager.compute()
    assert set(result) == {"val - prediction - accuracy", "val - aux - accuracy"}
    manager.clear()
    manager.update({"prediction": logits[0], "aux": logits[0]}, {"prediction": targets[0], "aux": targets[0]})
    assert manager.compute()["val - prediction - accuracy"] == manager.compute()["val - aux - accuracy"]
    with pytest.raises(AssertionError, match="keys of the targets do not match"):
        manager.update({"prediction": logits[0], "aux": logits[0]}, {"prediction": targets[0], "other": targets[0]})
    manager.reset()
    assert manager.metrics_per_prediction_type == {}


def test_loss_meter_averages_a_key_over_the_steps_that_reported_it() -> None:
    """An optional term (a penalty that is only active in some steps) is averaged over the steps that had it, as with the
    reference's list of per-step dictionaries; after ``clear`` a key that is not reported again does not linger."""
    import torch

    from fl4health_b200.utils.losses import LossMeter, LossMeterType, TrainingLosses

    meter = LossMeter(LossMeterType.AVERAGE, TrainingLosses)
    meter.update(TrainingLosses(torch.tensor(1.0), {"penalty": torch.tensor(4.0)}))
    meter.update(TrainingLosses(torch.tensor(3.0)))
    meter.update(TrainingLosses(torch.tensor(5.0), {"penalty": torch.tensor(2.0)}))
    assert meter.compute().as_dict() == {"penalty": 3.0, "backward": 3.0}
    meter.clear()
    meter.update(TrainingLosses(torch.tensor(2.0)))
    assert meter.compute().as_dict() == {"backward": 2.0}
    # the split form used around captured graphs: accumulate on the device, count on the host
    meter.clear()
    step = TrainingLosses(torch.tensor(1.0), {"penalty": torch.tensor(1.0)})
    for _ in range(4):
        meter.accumulate(step)
        meter.mark_step(losses=step)
    assert meter.compute().as_dict() == {"penalty": 1.0, "backward": 1.0} and meter.count == 4
    summed = LossMeter.aggregate_losses_dict([{"a": torch.tensor(1.0)}, {"a": torch.tensor(3.0), "b": torch.tensor(5.0)}], LossMeterType.AVERAGE)
    assert float(summed["a"]) == 2.0 and float(summed["b"]) == 5.0
