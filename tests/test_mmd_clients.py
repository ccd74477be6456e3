"""Ditto / MR-MTL with MK-MMD and Deep-MMD feature penalties: end-to-end federations on the in-process transport."""

from pathlib import Path

import pytest
import torch

from fl4health_b200.clients.deep_mmd_clients import DittoDeepMmdClient, MrMtlDeepMmdClient
from fl4health_b200.clients.ditto_client import DittoClient
from fl4health_b200.clients.mkmmd_clients import DittoMkMmdClient, MrMtlMkMmdClient
from fl4health_b200.metrics import Accuracy
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
from fl4health_b200.servers.adaptive_constraint_servers.ditto_server import DittoServer
from fl4health_b200.servers.adaptive_constraint_servers.mrmtl_server import MrMtlServer
from fl4health_b200.servers.client_manager import SimpleClientManager
from fl4health_b200.simulation import run_simulation
from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import SyntheticCifarMixin, TinyNet, fit_config_fn


def _federate(client_cls, server_cls, **client_kwargs):
    set_all_random_seeds(31)
    is_ditto = issubclass(client_cls, DittoClient)

    class Client(SyntheticCifarMixin, client_cls):
        model_fn = staticmethod(TinyNet)

        def get_optimizer(self, config):
            if is_ditto:
                return {"global": torch.optim.SGD(self.global_model.parameters(), lr=0.05),
                        "local": torch.optim.SGD(self.model.parameters(), lr=0.05)}
            return torch.optim.SGD(self.model.parameters(), lr=0.05)

    clients = []
    for idx in range(2):
        client = Client(Path("."), [Accuracy()], torch.device("cpu"), client_name=f"m{idx}", **client_kwargs)
        client.seed = idx
        clients.append(client)
    cfg = fit_config_fn(local_steps=4)
    strategy = FedAvgWithAdaptiveConstraint(
        initial_parameters=None, initial_loss_weight=0.1, min_fit_clients=2, min_evaluate_clients=2, min_available_clients=2,
        on_fit_config_fn=cfg, on_evaluate_config_fn=cfg, fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
        evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
    )
    server = server_cls(SimpleClientManager(), {"n_server_rounds": 2}, strategy, on_init_parameters_config_fn=cfg)
    history = run_simulation(server, clients, 2)
    return clients, history


@pytest.mark.parametrize("client_cls,server_cls", [(DittoMkMmdClient, DittoServer), (MrMtlMkMmdClient, MrMtlServer)])
@pytest.mark.parametrize("interval", [2, -1])
def test_mkmmd_clients(client_cls, server_cls, interval: int) -> None:
    clients, history = _federate(client_cls, server_cls, mkmmd_loss_weight=1.0, feature_extraction_layers=["bn"],
                                 beta_global_update_interval=interval, num_accumulating_batches=2)
    assert len(history.losses_distributed) == 2
    client = clients[0]
    loss = client.mkmmd_losses["bn"]
    assert loss.betas.sum().item() == pytest.approx(1.0, abs=1e-4) and bool((loss.betas >= 0).all())
    assert not client.engine.cuda_graphs and len(client.local_feature_extractor.fhooks) == 1
    fit_metrics = history.metrics_distributed_fit
    assert fit_metrics  # training ran and reported
    # the anchor is a frozen copy: no parameter of it requires grad
    assert all(not p.requires_grad for p in client.initial_global_model.parameters())


@pytest.mark.parametrize("client_cls,server_cls", [(DittoDeepMmdClient, DittoServer), (MrMtlDeepMmdClient, MrMtlServer)])
def test_deep_mmd_clients(client_cls, server_cls) -> None:
    clients, history = _federate(client_cls, server_cls, deep_mmd_loss_weight=1.0,
                                 feature_extraction_layers_with_size={"bn": 4 * 32 * 32}, mmd_kernel_train_interval=2,
                                 num_accumulating_batches=1)
    assert len(history.losses_distributed) == 2
    client = clients[0]
    kernel = client.deep_mmd_losses["bn"]
    assert not kernel.training  # left in eval mode after validation
    assert all(torch.isfinite(p).all() for p in kernel.featurizer.parameters())
