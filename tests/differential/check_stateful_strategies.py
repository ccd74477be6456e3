"""FedDG-GA (generalisation-adjusted aggregation weights over rounds) and client-level DP FedAvgM (clipping-bound
adaptation, server momentum, weighted / unweighted noisy aggregates with the noise turned down to nothing) vs the
reference."""
import numpy as np
import torch

import fl4health.parameter_exchange.parameter_packer as ref_pack
import fl4health.strategies.client_dp_fedavgm as ref_dp
import fl4health.strategies.feddg_ga as ref_ga
import fl4health_b200.strategies.client_dp_fedavgm as my_dp
import fl4health_b200.strategies.feddg_ga as my_ga

import flwr.common as fc
import fl4health_b200.common.typing as mt

rng = np.random.default_rng(23)
agreed = 0
SHAPES = [(4, 3), (4,), (2, 4), (2,)]


class Proxy:
    def __init__(self, cid: str) -> None:
        self.cid = cid


def weights() -> list[np.ndarray]:
    return [rng.normal(size=s).astype(np.float32) for s in SHAPES]


def numpy_list(parameters, ours: bool):
    out = (mt if ours else fc).parameters_to_ndarrays(parameters)
    return [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a) for a in out]


def same_arrays(a, b, tol=1e-5) -> None:
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert np.asarray(x).shape == np.asarray(y).shape
        assert np.allclose(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), atol=tol, rtol=tol), np.abs(np.asarray(x) - np.asarray(y)).max()


# -- FedDG-GA -------------------------------------------------------------------------------------------------------
for metric_type, step in (("LOSS", 0.2), ("ACCURACY", 0.5)):
    theirs = ref_ga.FedDgGa(fairness_metric=ref_ga.FairnessMetric(getattr(ref_ga.FairnessMetricType, metric_type)), adjustment_weight_step_size=step)
    ours = my_ga.FedDgGa(fairness_metric=my_ga.FairnessMetric(getattr(my_ga.FairnessMetricType, metric_type)), adjustment_weight_step_size=step)
    for strategy in (theirs, ours):  # what configure_fit derives from the round config and the sampled cohort
        strategy.num_rounds, strategy.initial_adjustment_weight = 5, 1.0 / 3
    name = theirs.fairness_metric.metric_name
    assert name == ours.fairness_metric.metric_name and theirs.fairness_metric.signal == ours.fairness_metric.signal
    for server_round in range(1, 6):
        payloads = [weights() for _ in range(3)]
        local = [float(rng.random()) for _ in range(3)]
        after = [float(rng.random()) for _ in range(3)]
        if server_round == 4:
            after = [value + 0.3 for value in local]  # identical gaps: the weights must not move
        fit_metrics = [{name: value, "val - checkpoint": value} for value in local]
        eval_metrics = [{name: value, "val - checkpoint": value} for value in after]
        fit_ref = [(Proxy(f"c{i}"), fc.FitRes(fc.Status(fc.Code.OK, ""), fc.ndarrays_to_parameters(p), 10 * (i + 1), dict(m))) for i, (p, m) in enumerate(zip(payloads, fit_metrics))]
        fit_mine = [(Proxy(f"c{i}"), mt.FitRes(mt.Status(mt.Code.OK, ""), mt.ndarrays_to_parameters(p), 10 * (i + 1), dict(m))) for i, (p, m) in enumerate(zip(payloads, fit_metrics))]
        p_ref, _ = theirs.aggregate_fit(server_round, fit_ref, [])
        p_mine, _ = ours.aggregate_fit(server_round, fit_mine, [])
        same_arrays(numpy_list(p_ref, False), numpy_list(p_mine, True))
        ev_ref = [(Proxy(f"c{i}"), fc.EvaluateRes(fc.Status(fc.Code.OK, ""), v, 10 * (i + 1), dict(m))) for i, (v, m) in enumerate(zip(after, eval_metrics))]
        ev_mine = [(Proxy(f"c{i}"), mt.EvaluateRes(mt.Status(mt.Code.OK, ""), v, 10 * (i + 1), dict(m))) for i, (v, m) in enumerate(zip(after, eval_metrics))]
        (l_ref, _), (l_mine, _) = theirs.aggregate_evaluate(server_round, ev_ref, []), ours.aggregate_evaluate(server_round, ev_mine, [])
        assert abs(l_ref - l_mine) < 1e-6
        assert theirs.adjustment_weights.keys() == ours.adjustment_weights.keys()
        for cid in theirs.adjustment_weights:
            assert abs(theirs.adjustment_weights[cid] - ours.adjustment_weights[cid]) < 1e-9, (server_round, theirs.adjustment_weights, ours.adjustment_weights)
        assert abs(theirs.get_current_weight_step_size(server_round) - ours.get_current_weight_step_size(server_round)) < 1e-12
    agreed += 1

# -- client-level DP FedAvgM ------------------------------------------------------------------------------------------
counts = [20, 30, 50]
for weighted, adaptive, beta, server_lr in ((False, False, 0.0, 1.0), (True, False, 0.9, 0.5), (False, True, 0.9, 1.0), (True, True, 0.5, 0.7)):
    initial = weights()
    noise = 1e-7 if adaptive else 0.0  # adaptive clipping divides by the noise multipliers: tiny instead of zero
    kwargs = dict(weighted_aggregation=weighted, adaptive_clipping=adaptive, server_learning_rate=server_lr, clipping_learning_rate=0.3,
                  clipping_quantile=0.6, initial_clipping_bound=0.4, weight_noise_multiplier=noise, clipping_noise_multiplier=noise,
                  beta=beta, fraction_fit=1.0, min_available_clients=3)
    theirs = ref_dp.ClientLevelDPFedAvgM(initial_parameters=fc.ndarrays_to_parameters([w.copy() for w in initial]), **kwargs)
    ours = my_dp.ClientLevelDPFedAvgM(initial_parameters=mt.ndarrays_to_parameters([w.copy() for w in initial]), **kwargs)
    theirs.sample_counts, ours.sample_counts = list(counts), list(counts)
    # the initial parameters now carry the clipping bound at the end
    same_arrays(numpy_list(theirs.initial_parameters, False), numpy_list(ours.initial_parameters, True))
    for server_round in range(1, 5):
        payloads = [ref_pack.ParameterPackerWithClippingBit().pack_parameters([0.05 * w for w in weights()], float(rng.random() > 0.5)) for _ in counts]
        fit_ref = [(Proxy(f"c{i}"), fc.FitRes(fc.Status(fc.Code.OK, ""), fc.ndarrays_to_parameters(p), n, {})) for i, (p, n) in enumerate(zip(payloads, counts))]
        fit_mine = [(Proxy(f"c{i}"), mt.FitRes(mt.Status(mt.Code.OK, ""), mt.ndarrays_to_parameters(p), n, {})) for i, (p, n) in enumerate(zip(payloads, counts))]
        p_ref, _ = theirs.aggregate_fit(server_round, fit_ref, [])
        p_mine, _ = ours.aggregate_fit(server_round, fit_mine, [])
        same_arrays(numpy_list(p_ref, False), numpy_list(p_mine, True), tol=1e-5)
        assert abs(theirs.clipping_bound - ours.clipping_bound) < 1e-6, (theirs.clipping_bound, ours.clipping_bound)
    agreed += 1
print("configs agree:", agreed)
