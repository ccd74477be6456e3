"""Round loop of a Flower-style server: configure -> fan out over a thread pool -> aggregate."""

from __future__ import annotations

import concurrent.futures
import timeit
from logging import INFO, WARNING

from ..common.logger import log
from ..common.typing import (
    Code,
    DisconnectRes,
    EvaluateIns,
    EvaluateRes,
    FitIns,
    FitRes,
    GetParametersIns,
    Parameters,
    ReconnectIns,
    Scalar,
)
from .client_manager import ClientManager
from .client_proxy import ClientProxy
from .history import History
from .strategy import FedAvg, Strategy

FitResultsAndFailures = tuple[list[tuple[ClientProxy, FitRes]], list[tuple[ClientProxy, FitRes] | BaseException]]
EvaluateResultsAndFailures = tuple[list[tuple[ClientProxy, EvaluateRes]], list[tuple[ClientProxy, EvaluateRes] | BaseException]]
ReconnectResultsAndFailures = tuple[list[tuple[ClientProxy, DisconnectRes]], list[tuple[ClientProxy, DisconnectRes] | BaseException]]


def _fan_out(verb: str, instructions: list, max_workers: int | None, timeout: float | None, group_id: int | None,
             needs_ok: bool = True) -> tuple[list, list]:
    """Issue ``proxy.<verb>(ins)`` for every (proxy, ins) pair concurrently; split outcomes into results / failures."""

    def one(proxy: ClientProxy, ins: object) -> tuple[ClientProxy, object]:
        return proxy, getattr(proxy, verb)(ins, timeout=timeout, group_id=group_id)

    with concurrent.futures.ThreadPoolExecutor(max_workers=max_workers) as pool:
        futures = {pool.submit(one, proxy, ins) for proxy, ins in instructions}
        done, _ = concurrent.futures.wait(futures, timeout=None)
    results: list = []
    failures: list = []
    for future in done:
        error = future.exception()
        if error is not None:
            failures.append(error)
            continue
        outcome = future.result()
        status = getattr(outcome[1], "status", None)
        if needs_ok and status is not None and status.code != Code.OK:
            failures.append(outcome)
        else:
            results.append(outcome)
    return results, failures


def fit_clients(client_instructions: list[tuple[ClientProxy, FitIns]], max_workers: int | None, timeout: float | None,
                group_id: int) -> FitResultsAndFailures:
    return _fan_out("fit", client_instructions, max_workers, timeout, group_id)


def evaluate_clients(client_instructions: list[tuple[ClientProxy, EvaluateIns]], max_workers: int | None,
                     timeout: float | None, group_id: int) -> EvaluateResultsAndFailures:
    return _fan_out("evaluate", client_instructions, max_workers, timeout, group_id)


def reconnect_clients(client_instructions: list[tuple[ClientProxy, ReconnectIns]], max_workers: int | None,
                      timeout: float | None) -> ReconnectResultsAndFailures:
    return _fan_out("reconnect", client_instructions, max_workers, timeout, None, needs_ok=False)


class Server:
    def __init__(self, *, client_manager: ClientManager, strategy: Strategy | None = None) -> None:
        self._client_manager = client_manager
        self.parameters = Parameters(tensors=[], tensor_type="numpy.ndarray")
        self.strategy: Strategy = strategy if strategy is not None else FedAvg()
        self.max_workers: int | None = None

    def set_max_workers(self, max_workers: int | None) -> None:
        self.max_workers = max_workers

    def set_strategy(self, strategy: Strategy) -> None:
        self.strategy = strategy

    def client_manager(self) -> ClientManager:
        return self._client_manager

    # ------------------------------------------------------------------ training loop
    def fit(self, num_rounds: int, timeout: float | None) -> tuple[History, float]:
        history = History()
        log(INFO, "[INIT]")
        self.parameters = self._get_initial_parameters(server_round=0, timeout=timeout)
        central = self.strategy.evaluate(0, parameters=self.parameters)
        if central is not None:
            history.add_loss_centralized(server_round=0, loss=central[0])
            history.add_metrics_centralized(server_round=0, metrics=central[1])
        start = timeit.default_timer()
        for current_round in range(1, num_rounds + 1):
            log(INFO, "")
            log(INFO, "[ROUND %s]", current_round)
            fitted = self.fit_round(server_round=current_round, timeout=timeout)
            if fitted is not None:
                new_parameters, fit_metrics, _ = fitted
                if new_parameters:
                    self.parameters = new_parameters
                history.add_metrics_distributed_fit(server_round=current_round, metrics=fit_metrics)
            central = self.strategy.evaluate(current_round, parameters=self.parameters)
            if central is not None:
                history.add_loss_centralized(server_round=current_round, loss=central[0])
                history.add_metrics_centralized(server_round=current_round, metrics=central[1])
            evaluated = self.evaluate_round(server_round=current_round, timeout=timeout)
            if evaluated is not None:
                loss, metrics, _ = evaluated
                if loss is not None:
                    history.add_loss_distributed(server_round=current_round, loss=loss)
                    history.add_metrics_distributed(server_round=current_round, metrics=metrics)
        return history, timeit.default_timer() - start

    def fit_round(self, server_round: int, timeout: float | None) -> tuple[Parameters | None, dict[str, Scalar], FitResultsAndFailures] | None:
        instructions = self.strategy.configure_fit(server_round=server_round, parameters=self.parameters,
                                                   client_manager=self._client_manager)
        if not instructions:
            log(INFO, "configure_fit: no clients selected, cancel")
            return None
        results, failures = fit_clients(instructions, self.max_workers, timeout, group_id=server_round)
        parameters, metrics = self.strategy.aggregate_fit(server_round, results, failures)
        return parameters, metrics, (results, failures)

    def evaluate_round(self, server_round: int, timeout: float | None) -> tuple[float | None, dict[str, Scalar], EvaluateResultsAndFailures] | None:
        instructions = self.strategy.configure_evaluate(server_round=server_round, parameters=self.parameters,
                                                        client_manager=self._client_manager)
        if not instructions:
            log(INFO, "configure_evaluate: no clients selected, skipping evaluation")
            return None
        results, failures = evaluate_clients(instructions, self.max_workers, timeout, group_id=server_round)
        loss, metrics = self.strategy.aggregate_evaluate(server_round, results, failures)
        return loss, metrics, (results, failures)

    def disconnect_all_clients(self, timeout: float | None) -> None:
        proxies = list(self._client_manager.all().values())
        reconnect_clients([(proxy, ReconnectIns(seconds=None)) for proxy in proxies], self.max_workers, timeout)

    def _get_initial_parameters(self, server_round: int, timeout: float | None) -> Parameters:
        parameters = self.strategy.initialize_parameters(client_manager=self._client_manager)
        if parameters is not None:
            return parameters
        chosen = self._client_manager.sample(1)[0]
        res = chosen.get_parameters(ins=GetParametersIns(config={}), timeout=timeout, group_id=server_round)
        if res.status.code != Code.OK:
            log(WARNING, "Failed to receive initial parameters from the client. Empty initial parameters will be used.")
        return res.parameters
