"""The generic federated server loop (role of ``flwr.server.server.Server``)."""

from __future__ import annotations

import timeit
from logging import INFO, WARNING
from typing import Any

from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.utils import tracing
from fl4health_b200.common.typing import Code, GetParametersIns, Parameters, ReconnectIns, Scalar
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.transport import (
    EvaluateResultsAndFailures,
    FitResultsAndFailures,
    LocalTransport,
)


class Server:
    def __init__(self, *, client_manager: ClientManager, strategy: Any = None, transport: Any = None) -> None:
        self._client_manager = client_manager
        self.parameters: Parameters = Parameters(tensors=[], tensor_type="torch")
        self.strategy = strategy
        self.max_workers: int | None = None
        self.transport = transport if transport is not None else LocalTransport()

    def set_max_workers(self, max_workers: int | None) -> None:
        self.max_workers = max_workers

    def set_strategy(self, strategy: Any) -> None:
        self.strategy = strategy

    def client_manager(self) -> ClientManager:
        return self._client_manager

    # ------------------------------------------------------------------------------------------------------
    def fit(self, num_rounds: int, timeout: float | None) -> tuple[History, float]:
        history = History()
        log(INFO, "[INIT]")
        self.parameters = self._get_initial_parameters(server_round=0, timeout=timeout)
        log(INFO, "Starting evaluation of initial global parameters")
        res = self.strategy.evaluate(0, parameters=self.parameters)
        if res is not None:
            log(INFO, "initial parameters (loss, other metrics): %s, %s", res[0], res[1])
            history.add_loss_centralized(server_round=0, loss=res[0])
            history.add_metrics_centralized(server_round=0, metrics=res[1])
        else:
            log(INFO, "Evaluation returned no results (`None`)")

        start_time = timeit.default_timer()
        for current_round in range(1, num_rounds + 1):
            log(INFO, "")
            log(INFO, "[ROUND %s]", current_round)
            res_fit = self.fit_round(server_round=current_round, timeout=timeout)
            if res_fit is not None:
                parameters_prime, fit_metrics, _ = res_fit
                if parameters_prime:
                    self.parameters = parameters_prime
                history.add_metrics_distributed_fit(server_round=current_round, metrics=fit_metrics)

            res_cen = self.strategy.evaluate(current_round, parameters=self.parameters)
            if res_cen is not None:
                loss_cen, metrics_cen = res_cen
                log(INFO, "fit progress: (%s, %s, %s, %s)", current_round, loss_cen, metrics_cen,
                    timeit.default_timer() - start_time)
                history.add_loss_centralized(server_round=current_round, loss=loss_cen)
                history.add_metrics_centralized(server_round=current_round, metrics=metrics_cen)

            res_fed = self.evaluate_round(server_round=current_round, timeout=timeout)
            if res_fed is not None:
                loss_fed, evaluate_metrics_fed, _ = res_fed
                if loss_fed is not None:
                    history.add_loss_distributed(server_round=current_round, loss=loss_fed)
                    history.add_metrics_distributed(server_round=current_round, metrics=evaluate_metrics_fed)
        elapsed = timeit.default_timer() - start_time
        return history, elapsed

    def evaluate_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[float | None, dict[str, Scalar], EvaluateResultsAndFailures] | None:
        client_instructions = self.strategy.configure_evaluate(
            server_round=server_round, parameters=self.parameters, client_manager=self._client_manager
        )
        if not client_instructions:
            log(INFO, "configure_evaluate: no clients selected, skipping evaluation")
            return None
        results, failures = self.transport.evaluate_clients(
            client_instructions, max_workers=self.max_workers, timeout=timeout, group_id=server_round
        )
        log(INFO, "aggregate_evaluate: received %s results and %s failures", len(results), len(failures))
        loss_aggregated, metrics_aggregated = self.strategy.aggregate_evaluate(server_round, results, failures)
        return loss_aggregated, metrics_aggregated, (results, failures)

    def fit_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[Parameters | None, dict[str, Scalar], FitResultsAndFailures] | None:
        client_instructions = self.strategy.configure_fit(
            server_round=server_round, parameters=self.parameters, client_manager=self._client_manager
        )
        if not client_instructions:
            log(INFO, "configure_fit: no clients selected, cancel")
            return None
        results, failures = self.transport.fit_clients(
            client_instructions, max_workers=self.max_workers, timeout=timeout, group_id=server_round
        )
        log(INFO, "aggregate_fit: received %s results and %s failures", len(results), len(failures))
        with tracing.phase("aggregate"):
            parameters_aggregated, metrics_aggregated = self.strategy.aggregate_fit(server_round, results, failures)
        return parameters_aggregated, metrics_aggregated, (results, failures)

    def disconnect_all_clients(self, timeout: float | None) -> None:
        for proxy in list(self._client_manager.all().values()):
            try:
                proxy.reconnect(ReconnectIns(seconds=None), timeout=timeout, group_id=None)
            except Exception as exc:  # noqa: BLE001
                log(WARNING, f"client {proxy.cid} failed to disconnect cleanly: {exc!r}")

    def _get_initial_parameters(self, server_round: int, timeout: float | None) -> Parameters:
        parameters = self.strategy.initialize_parameters(client_manager=self._client_manager)
        if parameters is not None:
            log(INFO, "Using initial global parameters provided by strategy")
            return parameters
        log(INFO, "Requesting initial parameters from one random client")
        random_client = self._client_manager.sample(1)[0]
        res = random_client.get_parameters(GetParametersIns(config={}), timeout=timeout, group_id=server_round)
        if res.status.code != Code.OK:
            log(WARNING, "Failed to receive initial parameters from the client. Empty initial parameters will be used.")
        return res.parameters
