"""Server-side aggregation of client metric dicts (parity: ``fl4health/metrics/metric_aggregation.py:6-172``)."""

from __future__ import annotations

from collections import defaultdict

from fl4health_b200.common.typing import Metrics


def _accumulate(store: Metrics, key: str, value: object, scale: int = 1) -> None:
    if isinstance(value, bool) or not isinstance(value, (int, float)):
        raise ValueError("Metric type is not supported")
    zero = 0.0 if isinstance(value, float) else 0
    store[key] = store.get(key, zero) + scale * value  # type: ignore[operator]


def uniform_metric_aggregation(
    all_client_metrics: list[tuple[int, Metrics]],
) -> tuple[defaultdict[str, int], Metrics]:
    """How many clients reported each metric, and the un-normalised *sums* over those clients (the counterpart of
    ``metric_aggregation``; ``uniform_normalize_metrics`` turns the pair into means)."""
    sums: Metrics = {}
    counts: defaultdict[str, int] = defaultdict(int)
    for _, client_metrics in all_client_metrics:
        for key, value in client_metrics.items():
            _accumulate(sums, key, value)
            counts[key] += 1
    return counts, sums


def metric_aggregation(all_client_metrics: list[tuple[int, Metrics]]) -> tuple[int, Metrics]:
    """Sample-weighted *sums* (un-normalized) plus the total number of examples."""
    sums: Metrics = {}
    total_examples = 0
    for num_examples, client_metrics in all_client_metrics:
        total_examples += num_examples
        for key, value in client_metrics.items():
            _accumulate(sums, key, value, num_examples)
    return total_examples, sums


def normalize_metrics(total_examples: int, aggregated_metrics: Metrics) -> Metrics:
    return {k: v / total_examples for k, v in aggregated_metrics.items() if isinstance(v, (int, float))}


def uniform_normalize_metrics(total_client_count_by_metric: defaultdict[str, int], aggregated_metrics: Metrics) -> Metrics:
    return {
        k: v / total_client_count_by_metric[k] for k, v in aggregated_metrics.items() if isinstance(v, (int, float))
    }


def fit_metrics_aggregation_fn(all_client_metrics: list[tuple[int, Metrics]]) -> Metrics:
    total_examples, sums = metric_aggregation(all_client_metrics)
    return normalize_metrics(total_examples, sums)


def evaluate_metrics_aggregation_fn(all_client_metrics: list[tuple[int, Metrics]]) -> Metrics:
    total_examples, sums = metric_aggregation(all_client_metrics)
    return normalize_metrics(total_examples, sums)


def uniform_evaluate_metrics_aggregation_fn(all_client_metrics: list[tuple[int, Metrics]]) -> Metrics:
    return uniform_normalize_metrics(*uniform_metric_aggregation(all_client_metrics))
