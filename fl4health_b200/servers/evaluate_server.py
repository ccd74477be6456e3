"""Federated-evaluation-only server (parity: ``fl4health/servers/evaluate_server.py``): no strategy, one evaluation pass;
optionally ships a checkpointed global model to the clients."""

from __future__ import annotations

import datetime
from collections.abc import Sequence
from logging import INFO, WARNING
from pathlib import Path
from typing import Any

import torch

from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import EvaluateIns, EvaluateRes, MetricsAggregationFn, Parameters, Scalar, ndarrays_to_parameters
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.servers.server import Server
from fl4health_b200.utils.random import generate_hash


class EvaluateServer(Server):
    def __init__(
        self,
        client_manager: ClientManager,
        fraction_evaluate: float,
        model_checkpoint_path: Path | None = None,
        evaluate_config: dict[str, Scalar] | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        accept_failures: bool = True,
        min_available_clients: int = 1,
        reporters: Sequence[BaseReporter] | None = None,
        transport: Any = None,
    ) -> None:
        super().__init__(client_manager=client_manager, strategy=None, transport=transport)
        self.model_checkpoint_path = model_checkpoint_path
        if model_checkpoint_path:
            self.parameters = self.load_model_checkpoint_to_parameters()
        self.fraction_evaluate = fraction_evaluate
        self.evaluate_config = evaluate_config
        self.min_available_clients = min_available_clients
        self.accept_failures = accept_failures
        self.evaluate_metrics_aggregation_fn = evaluate_metrics_aggregation_fn
        if fraction_evaluate < 1.0:
            log(INFO, f"Fraction Evaluate is {fraction_evaluate}. Thus, some clients may not participate in evaluation")
        self.server_name = generate_hash()
        self.reporters = list(reporters) if reporters is not None else []
        for reporter in self.reporters:
            reporter.initialize(id=self.server_name)

    def load_model_checkpoint_to_parameters(self) -> Parameters:
        assert self.model_checkpoint_path
        log(INFO, f"Loading model checkpoint at: {self.model_checkpoint_path}")
        model = torch.load(self.model_checkpoint_path, weights_only=False)
        return ndarrays_to_parameters([value.detach() for value in model.state_dict().values()])

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        history = History()
        log(INFO, "Federated Evaluation Starting")
        start = datetime.datetime.now()
        for reporter in self.reporters:
            reporter.report({"fit_start": str(start), "host_type": "server"})
        res_fed = self.federated_evaluate(timeout=timeout)
        end = datetime.datetime.now()
        for reporter in self.reporters:
            reporter.report({"fit_elapsed_time": str(end - start), "fit_end": str(end), "num_rounds": num_rounds, "host_type": "server"})
        if res_fed and res_fed[1]:
            history.add_metrics_distributed(server_round=0, metrics=res_fed[1])
            for reporter in self.reporters:
                reporter.report({"fit_metrics": res_fed[1]})
        log(INFO, "Federated Evaluation Finished in %s", str(end - start))
        return history, (end - start).total_seconds()

    def federated_evaluate(self, timeout: float | None) -> tuple[float | None, dict[str, Scalar], tuple[list, list]] | None:
        client_instructions = self.configure_evaluate()
        if not client_instructions:
            log(INFO, "Federated Evaluation: no clients selected, cancel")
            return None
        results, failures = self.transport.evaluate_clients(client_instructions, max_workers=self.max_workers, timeout=timeout, group_id=0)
        log(INFO, f"Federated Evaluation received {len(results)} results and {len(failures)} failures")
        _, metrics = self.aggregate_evaluate(results, failures)
        return None, metrics, (results, failures)

    def configure_evaluate(self) -> list[tuple[ClientProxy, EvaluateIns]]:
        if self.fraction_evaluate == 0.0:
            return []
        evaluate_ins = EvaluateIns(self.parameters, self.evaluate_config or {})
        if isinstance(self._client_manager, BaseFractionSamplingManager):
            clients = self._client_manager.sample_fraction(self.fraction_evaluate, self.min_available_clients)
        else:
            sample_size = int(self._client_manager.num_available() * self.fraction_evaluate)
            clients = self._client_manager.sample(num_clients=sample_size, min_num_clients=self.min_available_clients)
        return [(client, evaluate_ins) for client in clients]

    def aggregate_evaluate(self, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]) -> tuple[float | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        if self.evaluate_metrics_aggregation_fn:
            return None, self.evaluate_metrics_aggregation_fn([(res.num_examples, res.metrics) for _, res in results])
        log(WARNING, "No evaluate_metrics_aggregation_fn provided")
        return None, {}
