"""FedPM / FLASH / ensemble / partial exchange / evaluate-only / model-merge flows (CPU, in-process)."""

from pathlib import Path

import torch
from torch import nn

from fl4health_b200.checkpointing.checkpointer import LatestTorchModuleCheckpointer
from fl4health_b200.clients.ensemble_client import EnsembleClient
from fl4health_b200.clients.evaluate_client import EvaluateClient
from fl4health_b200.clients.fedpm_client import FedPmClient
from fl4health_b200.clients.flash_client import FlashClient
from fl4health_b200.clients.model_merge_client import ModelMergeClient
from fl4health_b200.clients.partial_weight_exchange_client import PartialWeightExchangeClient
from fl4health_b200.engine.data import BatchedTensorLoader
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.model_bases.ensemble_base import EnsembleModel
from fl4health_b200.model_bases.masked_layers.masked_layers_utils import is_masked_module
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.layer_exchanger import DynamicLayerExchanger
from fl4health_b200.parameter_exchange.parameter_selection_criteria import (
    LayerSelectionFunctionConstructor,
    largest_final_magnitude_scores,
)
from fl4health_b200.parameter_exchange.sparse_coo_parameter_exchanger import SparseCooParameterExchanger
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.servers.evaluate_server import EvaluateServer
from fl4health_b200.servers.fedpm_server import FedPmServer
from fl4health_b200.servers.model_merge_server import ModelMergeServer
from fl4health_b200.simulation import register_clients, run_simulation
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.fedavg_dynamic_layer import FedAvgDynamicLayer
from fl4health_b200.strategies.fedavg_sparse_coo_tensor import FedAvgSparseCooTensor
from fl4health_b200.strategies.fedpm import FedPm
from fl4health_b200.strategies.flash import Flash
from fl4health_b200.strategies.model_merge_strategy import ModelMergeStrategy
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import TinyNet, fit_config_fn, make_mixed_clients, synthetic_cifar


def _common(cfg=None):
    cfg = cfg or fit_config_fn()
    return dict(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2, on_fit_config_fn=cfg,
                on_evaluate_config_fn=cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)


def test_fedpm_end_to_end() -> None:
    set_all_random_seeds(51)

    def cfg(r):
        return {"current_server_round": r, "local_steps": 4, "batch_size": 32, "is_masked_model": False}

    clients = make_mixed_clients(FedPmClient, 2, model_fn=staticmethod(TinyNet), lr=0.5)
    server = FedPmServer(SimpleClientManager(), {"n_server_rounds": 3}, FedPm(**_common(cfg)), reset_frequency=2,
                         on_init_parameters_config_fn=cfg)
    history = run_simulation(server, clients, 3)
    assert len(history.losses_distributed) == 3
    model = clients[0].model
    assert any(is_masked_module(m) for m in model.modules())
    assert all(not p.requires_grad for n, p in model.named_parameters() if not n.endswith("_scores"))
    # both clients received identical score tensors (logit of the server's theta); frozen weights are never exchanged
    s0, s1 = clients[0].model.state_dict(), clients[1].model.state_dict()
    assert all(torch.allclose(s0[k], s1[k]) for k in s0 if k.endswith("_scores"))
    assert set(server.strategy.beta_parameters) == {k for k in s0 if k.endswith("_scores")}


def test_flash_client_and_strategy() -> None:
    set_all_random_seeds(52)

    def cfg(r):
        return {"current_server_round": r, "local_epochs": 3, "batch_size": 32, "gamma": 100.0}

    clients = make_mixed_clients(FlashClient, 2, n_train=64)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, Flash(initial_parameters=None, eta=0.05, **_common(cfg)),
                      on_init_parameters_config_fn=cfg)
    history = run_simulation(server, clients, 2)
    assert len(history.losses_distributed) == 2
    # gamma is huge: the first epoch can never stop (previous loss = inf), the second always does -> 2 epochs x 2 steps x 2 rounds
    assert clients[0].gamma == 100.0 and clients[0].total_steps == 8


def test_ensemble_client() -> None:
    set_all_random_seeds(53)
    clients = make_mixed_clients(EnsembleClient, 2, model_fn=staticmethod(
        lambda: EnsembleModel({"a": TinyNet(), "b": TinyNet()})))
    for c in clients:
        c.get_optimizer = (lambda self, config: {k: torch.optim.SGD(m.parameters(), lr=0.05)
                                                 for k, m in self.model.ensemble_models.items()}).__get__(c)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, BasicFedAvg(**_common()),
                      on_init_parameters_config_fn=fit_config_fn())
    history = run_simulation(server, clients, 2)
    assert "val - ensemble-pred - accuracy" in history.metrics_distributed
    assert {"val - a - accuracy", "val - b - accuracy"} <= set(history.metrics_distributed)


def test_dynamic_layer_and_sparse_partial_exchange() -> None:
    for kind in ("dynamic", "sparse"):
        set_all_random_seeds(54)
        clients = make_mixed_clients(PartialWeightExchangeClient, 2, model_fn=staticmethod(TinyNet))
        for c in clients:
            c.store_initial_model = True
            if kind == "dynamic":
                selector = LayerSelectionFunctionConstructor(1e-9, 0.5, normalize=False).select_by_percentage()
                c.get_parameter_exchanger = lambda config, s=selector: DynamicLayerExchanger(s)
            else:
                c.get_parameter_exchanger = lambda config: SparseCooParameterExchanger(0.3, largest_final_magnitude_scores)
        strategy = FedAvgDynamicLayer(**_common()) if kind == "dynamic" else FedAvgSparseCooTensor(**_common())
        server = FlServer(SimpleClientManager(), {"n_server_rounds": 2}, strategy, on_init_parameters_config_fn=fit_config_fn())
        history = run_simulation(server, clients, 2)
        assert len(history.losses_distributed) == 2, kind


def test_federated_evaluation_only(tmp_path: Path) -> None:
    torch.manual_seed(0)
    global_model = TinyNet()
    torch.save(global_model, tmp_path / "global.pkl")
    torch.save(TinyNet(), tmp_path / "local.pkl")

    class Client(EvaluateClient):
        def get_data_loader(self, config):
            return (BatchedTensorLoader(synthetic_cifar(64, 3), 32),)

        def get_criterion(self, config):
            return nn.CrossEntropyLoss()

        def initialize_global_model(self, config):
            return TinyNet()

    clients = [Client(Path("."), [Accuracy()], torch.device("cpu"), model_checkpoint_path=tmp_path / "local.pkl",
                      client_name=f"e{i}") for i in range(2)]
    server = EvaluateServer(SimpleClientManager(), fraction_evaluate=1.0, model_checkpoint_path=tmp_path / "global.pkl",
                            evaluate_config={"current_server_round": 0}, min_available_clients=2,
                            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    register_clients(server, clients)
    history, _ = server.fit(1)
    metrics = history.metrics_distributed
    assert "global_eval_manager - prediction - accuracy" in metrics and "local_eval_manager - prediction - accuracy" in metrics
    sd = clients[0].global_model.state_dict()
    assert all(torch.allclose(sd[k], v) for k, v in global_model.state_dict().items())


def test_model_merge(tmp_path: Path) -> None:
    class Client(ModelMergeClient):
        def get_model(self, config):
            torch.manual_seed(int(self.client_name[-1]))
            return TinyNet()

        def get_test_data_loader(self, config):
            return BatchedTensorLoader(synthetic_cifar(64, 9), 32)

    clients = [Client(Path("."), Path("."), [Accuracy()], torch.device("cpu"), client_name=f"m{i}") for i in range(2)]
    strategy = ModelMergeStrategy(min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2,
                                  fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                                  evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, weighted_aggregation=False)
    server = ModelMergeServer(SimpleClientManager(), strategy, LatestTorchModuleCheckpointer(str(tmp_path), "merged.pkl"),
                              TinyNet(), FullParameterExchanger())
    register_clients(server, clients)
    expected = {}
    for i in range(2):
        torch.manual_seed(i)
        for k, v in TinyNet().state_dict().items():
            expected[k] = expected.get(k, 0) + v.double() / 2
    history, _ = server.fit(1)
    merged = torch.load(tmp_path / "merged.pkl", weights_only=False).state_dict()
    for k, v in expected.items():
        if v.is_floating_point():
            assert torch.allclose(merged[k].double(), v, atol=1e-6), k
    assert "test - predictions - accuracy" in history.metrics_distributed
