"""Batch-accumulating metrics, their wrappers and the federated metric aggregation functions vs the reference."""
import math

import torch

import fl4health.metrics.compound_metrics as ref_compound
import fl4health.metrics.metric_aggregation as ref_agg
import fl4health.metrics.metrics as ref
import fl4health_b200.metrics.compound_metrics as my_compound
import fl4health_b200.metrics.metric_aggregation as my_agg
import fl4health_b200.metrics.metrics as mine

torch.manual_seed(3)
agreed = 0


def same(a: dict, b: dict, tol: float = 1e-6) -> None:
    assert a.keys() == b.keys(), (a, b)
    for key in a:
        assert math.isclose(float(a[key]), float(b[key]), rel_tol=tol, abs_tol=tol), (key, a[key], b[key])


def batches(kind: str):
    for _ in range(4):
        if kind == "multiclass":
            yield torch.randn(16, 5), torch.randint(0, 5, (16,))
        elif kind == "binary_logit":
            target = torch.randint(0, 2, (16, 1))
            yield torch.rand(16, 1) * 0.8 + 0.1 * target, target
        elif kind == "binary_prob2":
            target = torch.randint(0, 2, (16,))
            yield torch.softmax(torch.randn(16, 2) + 2 * torch.nn.functional.one_hot(target, 2), dim=1), target
        elif kind == "volumes":  # the default spatial axes are (2, 3, 4)
            yield torch.rand(3, 1, 4, 6, 6), (torch.rand(3, 1, 4, 6, 6) > 0.6).float()
        else:  # images
            yield torch.rand(4, 1, 8, 8), (torch.rand(4, 1, 8, 8) > 0.6).float()


cases = [
    ("Accuracy", {}, "multiclass"), ("Accuracy", {}, "binary_logit"), ("BalancedAccuracy", {}, "multiclass"),
    ("F1", {}, "multiclass"), ("F1", {"average": "macro"}, "multiclass"), ("F1", {"average": "micro"}, "multiclass"),
    ("RocAuc", {}, "multiclass_prob"),  # (the reference hands 2-column scores to sklearn, which rejects them: multiclass only)
    ("BinarySoftDiceCoefficient", {}, "volumes"), ("BinarySoftDiceCoefficient", {"logits_threshold": None}, "volumes"),
    ("BinarySoftDiceCoefficient", {"spatial_dimensions": (1, 2, 3), "epsilon": 1e-3}, "images"),
]
for name, kwargs, kind in cases:
    theirs, ours = getattr(ref, name)(**kwargs), getattr(mine, name)(**kwargs)
    data = list(batches(kind)) if kind != "multiclass_prob" else [
        (torch.softmax(torch.randn(40, 4), dim=1), torch.arange(40) % 4) for _ in range(3)]
    for pred, target in data:
        theirs.update(pred, target)
        ours.update(pred, target)
        same(theirs.compute("val"), ours.compute("val"), 1e-5)
    same(theirs.compute(), ours.compute(), 1e-5)
    theirs.clear(); ours.clear()
    pred, target = data[0]
    theirs.update(pred, target); ours.update(pred, target)
    same(theirs.compute("x"), ours.compute("x"), 1e-5)
    agreed += 1

# wrappers: exponential moving average over rounds, transforms before the metric
for smoothing in (0.1, 0.5):
    theirs = ref_compound.EmaMetric(ref.Accuracy(), smoothing_factor=smoothing, name="ema")
    ours = my_compound.EmaMetric(mine.Accuracy(), smoothing_factor=smoothing, name="ema")
    for _ in range(4):  # four "rounds": update, compute (which advances the average), clear
        for pred, target in batches("multiclass"):
            theirs.update(pred, target); ours.update(pred, target)
        same(theirs.compute("r"), ours.compute("r"))
        theirs.clear(); ours.clear()
    agreed += 1
squash = [torch.sigmoid, lambda x: (x > 0.5).float()]
theirs = ref_compound.TransformsMetric(ref.Accuracy(), pred_transforms=squash, target_transforms=[lambda t: t.long()])
ours = my_compound.TransformsMetric(mine.Accuracy(), pred_transforms=squash, target_transforms=[lambda t: t.long()])
for _ in range(3):
    pred, target = torch.randn(12, 1), torch.randint(0, 2, (12, 1)).float()
    theirs.update(pred, target); ours.update(pred, target)
same(theirs.compute("t"), ours.compute("t"))
agreed += 1

# server-side aggregation of client metrics
reports = [(10, {"acc": 0.5, "loss": 1.0, "note": "a"}), (30, {"acc": 0.9, "loss": 0.2, "note": "b"}), (60, {"acc": 0.7, "loss": 0.4, "note": "c"})]
numeric = [(n, {k: v for k, v in m.items() if k != "note"}) for n, m in reports]
for fn in ("fit_metrics_aggregation_fn", "evaluate_metrics_aggregation_fn", "uniform_evaluate_metrics_aggregation_fn"):
    same(getattr(ref_agg, fn)(numeric), getattr(my_agg, fn)(numeric)); agreed += 1
a, b = ref_agg.metric_aggregation(numeric), my_agg.metric_aggregation(numeric)
assert a[0] == b[0]; same(a[1], b[1]); agreed += 1
a, b = ref_agg.uniform_metric_aggregation(numeric), my_agg.uniform_metric_aggregation(numeric)
assert dict(a[0]) == dict(b[0]); same(a[1], b[1]); agreed += 1
same(ref_agg.normalize_metrics(100, {"acc": 70.0}), my_agg.normalize_metrics(100, {"acc": 70.0})); agreed += 1
print("configs agree:", agreed)
