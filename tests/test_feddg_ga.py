"""FedDG-GA oracles (values from the reference's tests/strategies/test_feddg_ga.py) + an APFL-backed e2e run."""

import pytest
import torch

from fl4health_b200.client_managers.fixed_sampling_client_manager import FixedSamplingClientManager
from fl4health_b200.clients.apfl_client import ApflClient
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.model_bases.apfl_base import ApflModule
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.feddg_ga import FairnessMetric, FairnessMetricType, FedDgGa
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import TinyNet, make_mixed_clients

LOSS = FairnessMetricType.LOSS.value


def _primed(eval0: float, eval1: float) -> FedDgGa:
    strategy = FedDgGa()
    strategy.num_rounds = 3
    strategy.initial_adjustment_weight = 1 / 3
    strategy.train_metrics = {"1": {LOSS: 0.5467}, "2": {LOSS: 0.5432}}
    strategy.evaluation_metrics = {"1": {LOSS: eval0}, "2": {LOSS: eval1}}
    strategy.adjustment_weights = {"1": 1 / 3, "2": 1 / 3}
    return strategy


def test_update_weights_by_ga() -> None:
    strategy = _primed(0.3556, 0.7654)
    strategy.update_weights_by_ga(2, ["1", "2"])
    assert strategy.adjustment_weights["1"] == pytest.approx(0.2999, abs=5e-4)
    assert strategy.adjustment_weights["2"] == pytest.approx(0.7000, abs=5e-4)
    same = _primed(0.5467, 0.5432)
    same.update_weights_by_ga(2, ["1", "2"])
    assert same.adjustment_weights == {"1": 0.5, "2": 0.5}


def test_step_size_schedule_and_metric_signals() -> None:
    strategy = FedDgGa()
    strategy.num_rounds = 3
    assert [strategy.get_current_weight_step_size(r) for r in (1, 2, 3)] == [
        pytest.approx(0.2, abs=5e-4), pytest.approx(0.1333, abs=5e-4), pytest.approx(0.0666, abs=5e-4)]
    assert FairnessMetric(FairnessMetricType.ACCURACY).signal == -1.0 and FairnessMetric(FairnessMetricType.LOSS).signal == 1.0
    with pytest.raises(AssertionError):
        FairnessMetric(FairnessMetricType.CUSTOM)
    with pytest.raises(AssertionError):
        FedDgGa(adjustment_weight_step_size=1.5)


def test_feddg_ga_end_to_end_over_apfl() -> None:
    set_all_random_seeds(71)

    def cfg(r):
        return {"current_server_round": r, "local_steps": 4, "batch_size": 32, "n_server_rounds": 3,
                "evaluate_after_fit": True, "pack_losses_with_val_metrics": True}

    clients = make_mixed_clients(ApflClient, 2, model_fn=staticmethod(lambda: ApflModule(TinyNet())))
    for c in clients:
        c.get_optimizer = (lambda self, config: {
            "local": torch.optim.SGD(self.model.local_model.parameters(), lr=0.05),
            "global": torch.optim.SGD(self.model.global_model.parameters(), lr=0.05)}).__get__(c)
    strategy = FedDgGa(on_fit_config_fn=cfg, on_evaluate_config_fn=cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                       evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    server = FlServer(FixedSamplingClientManager(), {"n_server_rounds": 3}, strategy, on_init_parameters_config_fn=cfg)
    history = run_simulation(server, clients, 2)  # (n_server_rounds=3 only sets the step-size schedule)
    assert len(history.losses_distributed) == 2
    assert set(strategy.adjustment_weights) == {"c0", "c1"}
    assert sum(strategy.adjustment_weights.values()) == pytest.approx(1.0)
    assert strategy.adjustment_weights["c0"] != 0.5
