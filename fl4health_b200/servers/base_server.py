"""``FlServer``: the federated round loop with reporting, checkpointing, resume and a failure policy.

API parity: ``fl4health/servers/base_server.py:36-643`` (constructor, ``fit``, ``fit_round``, ``evaluate_round``,
``poll_clients_for_sample_counts``, ``update_before_fit``, ``shutdown``, test/val metric split, report keys).
The loop is driven directly (no Flower, no gRPC): clients are reached through ``self.transport`` which is either the
in-process ``LocalTransport`` or the SPMD transport (one client per GPU rank, collectives for the payloads).
Device-timed per-round durations are reported next to the reference's integer-second wall-clock keys.
"""

from __future__ import annotations

import datetime
from collections.abc import Callable, Sequence
from logging import ERROR, INFO, WARNING
from typing import Any

import torch

from fl4health_b200.checkpointing.server_module import BaseServerCheckpointAndStateModule
from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Code, Config, EvaluateRes, GetParametersIns, Parameters, Scalar
from fl4health_b200.metrics.base_metrics import TEST_LOSS_KEY, TEST_NUM_EXAMPLES_KEY, MetricPrefix
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.servers.server import Server
from fl4health_b200.servers.transport import EvaluateResultsAndFailures, FitResultsAndFailures
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.strategies.strategy import Strategy
from fl4health_b200.strategies.strategy_with_poll import StrategyWithPolling
from fl4health_b200.utils.random import generate_hash
from fl4health_b200.utils.typing import EvaluateFailures, FitFailures


class FlServer(Server):
    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        strategy: Strategy | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        checkpoint_and_state_module: BaseServerCheckpointAndStateModule | None = None,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        transport: Any = None,
    ) -> None:
        super().__init__(client_manager=client_manager, strategy=strategy, transport=transport)
        self.fl_config = fl_config
        self.checkpoint_and_state_module = checkpoint_and_state_module or BaseServerCheckpointAndStateModule(
            model=None, parameter_exchanger=None, model_checkpointers=None, state_checkpointer=None
        )
        self.on_init_parameters_config_fn = on_init_parameters_config_fn
        self.server_name = server_name if server_name is not None else generate_hash()
        log(INFO, f"Server Name: {self.server_name}")
        self.accept_failures = accept_failures
        self.current_round: int
        self.history: History
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.server_name)
        # observability: callables invoked with the round number after each completed round (fit + evaluate)
        self.round_end_hooks: list[Callable[[int], None]] = []
        self._log_fl_config()

    # ------------------------------------------------------------------------------------------------------
    def update_before_fit(self, num_rounds: int, timeout: float | None) -> None:
        """Hook run once before the round loop (plan negotiation, warm starts...)."""

    def report_centralized_eval(self, history: History, num_rounds: int) -> None:
        if not history.losses_centralized:
            return
        by_round = dict(history.losses_centralized)
        for server_round in range(1, num_rounds + 1):
            if server_round not in by_round:
                continue
            self.reports_manager.report({"val - loss - centralized": by_round[server_round]}, server_round)
            round_metrics = {
                metric: dict(values)[server_round]
                for metric, values in history.metrics_centralized.items()
                if server_round in dict(values)
            }
            self.reports_manager.report({"eval_round_metrics_centralized": round_metrics}, server_round)

    def _run_rounds(self, num_rounds: int, timeout: float | None, resumable: bool) -> tuple[History, float]:
        log(INFO, "Initializing server state and global parameters")
        self.parameters = self._get_initial_parameters(server_round=0, timeout=timeout)
        self.history = History()
        self.current_round = 1
        if resumable:
            loaded = self._load_server_state()
            log(INFO, "Server state checkpoint successfully loaded." if loaded else "No server state checkpoint found. Starting from scratch.")
        if self.current_round == 1:
            log(INFO, "Evaluating initial parameters")
            res = self.strategy.evaluate(0, parameters=self.parameters)
            if res is not None:
                log(INFO, "initial parameters (loss, other metrics): %s, %s", res[0], res[1])
                self.history.add_loss_centralized(server_round=0, loss=res[0])
                self.history.add_metrics_centralized(server_round=0, metrics=res[1])
            log(INFO, "FL starting")

        start_time = datetime.datetime.now()
        while self.current_round < num_rounds + 1:
            log(INFO, "[ROUND %s]", self.current_round)
            res_fit = self.fit_round(server_round=self.current_round, timeout=timeout)
            if res_fit:
                parameters_prime, fit_metrics, _ = res_fit
                if parameters_prime:
                    self.parameters = parameters_prime
                self.history.add_metrics_distributed_fit(server_round=self.current_round, metrics=fit_metrics)

            res_cen = self.strategy.evaluate(self.current_round, parameters=self.parameters)
            if res_cen is not None:
                loss_cen, metrics_cen = res_cen
                log(INFO, "fit progress: (%s, %s, %s, %s)", self.current_round, loss_cen, metrics_cen,
                    (datetime.datetime.now() - start_time).total_seconds())
                self.history.add_loss_centralized(server_round=self.current_round, loss=loss_cen)
                self.history.add_metrics_centralized(server_round=self.current_round, metrics=metrics_cen)

            res_fed = self.evaluate_round(server_round=self.current_round, timeout=timeout)
            if res_fed:
                loss_fed, evaluate_metrics_fed, _ = res_fed
                if loss_fed is not None:
                    self.history.add_loss_distributed(server_round=self.current_round, loss=loss_fed)
                    self.history.add_metrics_distributed(server_round=self.current_round, metrics=evaluate_metrics_fed)

            for hook in self.round_end_hooks:
                hook(self.current_round)
            self.current_round += 1
            if resumable:
                self._save_server_state()
            self._maybe_inject_fault()

        elapsed = datetime.datetime.now() - start_time
        log(INFO, "FL finished in %s", str(elapsed))
        return self.history, elapsed.total_seconds()

    def fit_with_per_round_checkpointing(self, num_rounds: int, timeout: float | None) -> tuple[History, float]:
        """Resume-able loop: reloads ``current_round``/``history``/parameters if a state file exists and saves the
        server state after every round."""
        return self._run_rounds(num_rounds, timeout, resumable=True)

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        start_time = datetime.datetime.now()
        self.reports_manager.report({"fit_start": str(start_time), "host_type": "server"})
        self.update_before_fit(num_rounds, timeout)
        resumable = self.checkpoint_and_state_module.state_checkpointer is not None
        history, elapsed_time = self._run_rounds(num_rounds, timeout, resumable=resumable)
        end_time = datetime.datetime.now()
        self.reports_manager.report(
            {
                "fit_elapsed_time": round((end_time - start_time).total_seconds()),
                "fit_end": str(end_time),
                "num_rounds": num_rounds,
                "host_type": "server",
            }
        )
        self.report_centralized_eval(history, num_rounds)
        log(INFO, "[SUMMARY]\n%s", history)
        return history, elapsed_time

    def fit_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[Parameters | None, dict[str, Scalar], FitResultsAndFailures] | None:
        round_start = datetime.datetime.now()
        fit_round_results = super().fit_round(server_round, timeout)
        round_end = datetime.datetime.now()
        self.reports_manager.report(
            {
                "fit_round_start": str(round_start),
                "fit_round_end": str(round_end),
                "fit_round_time_elapsed": round((round_end - round_start).total_seconds()),
                "fit_round_seconds": (round_end - round_start).total_seconds(),
            },
            server_round,
        )
        if fit_round_results is not None:
            _, metrics, fit_results_and_failures = fit_round_results
            self.reports_manager.report({"fit_round_metrics": metrics}, server_round)
            failures = fit_results_and_failures[1] if fit_results_and_failures else None
            if failures and not self.accept_failures:
                self._log_client_failures(failures)
                self._terminate_after_unacceptable_failures(timeout)
        return fit_round_results

    def shutdown(self) -> None:
        self.reports_manager.report({"shutdown": str(datetime.datetime.now())})
        self.reports_manager.shutdown()

    def poll_clients_for_sample_counts(self, timeout: float | None) -> list[int]:
        log(INFO, "Polling Clients for sample counts")
        assert isinstance(self.strategy, StrategyWithPolling)
        client_instructions = self.strategy.configure_poll(server_round=1, client_manager=self._client_manager)
        results, _ = self.transport.poll_clients(client_instructions, max_workers=self.max_workers, timeout=timeout)
        sample_counts = [int(res.properties["num_train_samples"]) for _, res in results]
        log(INFO, f"Polling complete: Retrieved {len(sample_counts)} sample counts")
        return sample_counts

    def evaluate_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[float | None, dict[str, Scalar], EvaluateResultsAndFailures] | None:
        start_time = datetime.datetime.now()
        eval_round_results = self._evaluate_round(server_round, timeout)
        end_time = datetime.datetime.now()
        if eval_round_results:
            loss_aggregated, metrics_aggregated, (_, failures) = eval_round_results
            if failures and not self.accept_failures:
                self._log_client_failures(failures)
                self._terminate_after_unacceptable_failures(timeout)
            if loss_aggregated is not None:
                self._maybe_checkpoint(loss_aggregated, metrics_aggregated, server_round)
                report_data: dict[str, Any] = {
                    "val - loss - aggregated": loss_aggregated,
                    "round": server_round,
                    "eval_round_start": str(start_time),
                    "eval_round_end": str(end_time),
                    "eval_round_time_elapsed": round((end_time - start_time).total_seconds()),
                }
                if self.fl_config.get("local_epochs") is not None:
                    report_data["fit_epoch"] = server_round * self.fl_config["local_epochs"]  # type: ignore[operator]
                elif self.fl_config.get("local_steps") is not None:
                    report_data["fit_step"] = server_round * self.fl_config["local_steps"]  # type: ignore[operator]
                self.reports_manager.report(report_data, server_round)
                if metrics_aggregated:
                    self.reports_manager.report({"eval_round_metrics_aggregated": metrics_aggregated}, server_round)
        return eval_round_results

    # ------------------------------------------------------------------------------------------------------
    def _log_fl_config(self) -> None:
        log(INFO, "FL Configuration:" if self.fl_config else "FL Config is Empty")
        for key, value in self.fl_config.items():
            if not isinstance(value, bytes):
                log(INFO, f"Key: {key} Value: {value!r}")

    def _save_server_state(self) -> None:
        assert self.checkpoint_and_state_module.state_checkpointer is not None
        if self.transport.is_coordinator():
            self.checkpoint_and_state_module.save_state(self, self.parameters)

    def _load_server_state(self) -> bool:
        assert self.checkpoint_and_state_module.state_checkpointer is not None
        server_parameters = self.checkpoint_and_state_module.maybe_load_state(self)
        if server_parameters:
            self.parameters = server_parameters
            log(INFO, "Loaded server state from checkpoint")
            return True
        return False

    def _terminate_after_unacceptable_failures(self, timeout: float | None) -> None:
        assert not self.accept_failures
        self.disconnect_all_clients(timeout=timeout)
        self.shutdown()
        raise ValueError(
            f"The server encountered failures from the clients and accept_failures is set to {self.accept_failures}"
        )

    def _log_client_failures(self, failures: FitFailures | EvaluateFailures) -> None:
        log(ERROR, f"There were {len(failures)} failures in the fitting process. This will result in termination of the FL process")
        for failure in failures:
            if isinstance(failure, BaseException):
                log(ERROR, f"An exception was returned instead of any failed results: {failure!r}")
            else:
                client_proxy, _ = failure
                log(ERROR, f"Client {client_proxy.cid} failed but did not return an exception. Partial results were received")

    def _maybe_checkpoint(self, loss_aggregated: float, metrics_aggregated: dict[str, Scalar], server_round: int) -> None:
        if self.transport.is_coordinator():
            self.checkpoint_and_state_module.maybe_checkpoint(self.parameters, loss_aggregated, metrics_aggregated)

    def _maybe_inject_fault(self) -> None:
        """Fault injection for resume tests: ``FL4H_FAULT_AFTER_ROUND=r`` aborts the process after round r finished
        (and its state was saved), imitating a pre-emption."""
        import os

        target = os.environ.get("FL4H_FAULT_AFTER_ROUND")
        if target is not None and int(target) == self.current_round - 1:
            raise SystemExit(f"fault injected after round {target}")

    def _get_initial_parameters(self, server_round: int, timeout: float | None) -> Parameters:
        parameters = self.strategy.initialize_parameters(client_manager=self._client_manager)
        if parameters is not None:
            log(INFO, "Using initial global parameters provided by strategy")
            return parameters
        log(INFO, "Requesting initial parameters from one random client")
        if isinstance(self._client_manager, BaseFractionSamplingManager):
            random_client = self._client_manager.sample_one()[0]
        else:
            random_client = self._client_manager.sample(1)[0]
        if self.on_init_parameters_config_fn is None:
            log(WARNING, "on_init_parameters_config_fn is None; clients usually need a config to set themselves up.")
            ins = GetParametersIns(config={})
        else:
            ins = GetParametersIns(config=self.on_init_parameters_config_fn(server_round))
        res = self.transport_get_parameters(random_client, ins, timeout, server_round)
        if res.status.code == Code.OK:
            log(INFO, "Received initial parameters from one random client")
        else:
            log(WARNING, "Failed to receive initial parameters from the client. Empty initial parameters will be used.")
        initial_parameters = res.parameters
        # detach from the client's live arena: the global model must not alias a client's training buffer
        initial_parameters = Parameters(
            tensors=[t.detach().clone() if isinstance(t, torch.Tensor) else t for t in initial_parameters.tensors],
            tensor_type=initial_parameters.tensor_type,
        )
        if isinstance(self.strategy, BasicFedAvg):
            self.strategy.add_auxiliary_information(initial_parameters)
        return initial_parameters

    def transport_get_parameters(self, proxy: ClientProxy, ins: GetParametersIns, timeout: float | None, server_round: int) -> Any:
        getter = getattr(self.transport, "get_parameters", None)
        if getter is not None:
            return getter(proxy, ins, timeout, server_round)
        return proxy.get_parameters(ins=ins, timeout=timeout, group_id=server_round)

    # ------------------------------------------------------------------------------------------------------
    def _unpack_metrics(
        self, results: list[tuple[ClientProxy, EvaluateRes]]
    ) -> tuple[list[tuple[ClientProxy, EvaluateRes]], list[tuple[ClientProxy, EvaluateRes]]]:
        """Split each client's metrics into validation and ``"test - "``-prefixed parts (own loss / sample count)."""
        val_results, test_results = [], []
        test_prefix = MetricPrefix.TEST_PREFIX.value
        for proxy, res in results:
            val_metrics = {k: v for k, v in res.metrics.items() if not k.startswith(test_prefix)}
            test_metrics = {k: v for k, v in res.metrics.items() if k.startswith(test_prefix)}
            if test_metrics:
                assert TEST_LOSS_KEY in test_metrics and TEST_NUM_EXAMPLES_KEY in test_metrics, (
                    f"'{TEST_NUM_EXAMPLES_KEY}' and '{TEST_LOSS_KEY}' keys must be present in test_metrics for aggregation"
                )
                test_loss = float(test_metrics.pop(TEST_LOSS_KEY))  # type: ignore[arg-type]
                test_count = int(test_metrics.pop(TEST_NUM_EXAMPLES_KEY))  # type: ignore[arg-type]
                test_results.append((proxy, EvaluateRes(res.status, test_loss, test_count, test_metrics)))
            val_results.append((proxy, EvaluateRes(res.status, res.loss, res.num_examples, val_metrics)))
        return val_results, test_results

    def _handle_result_aggregation(
        self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]
    ) -> tuple[float | None, dict[str, Scalar]]:
        val_results, test_results = self._unpack_metrics(results)
        val_loss, val_metrics = self.strategy.aggregate_evaluate(server_round, val_results, failures)
        if test_results:
            test_loss, test_metrics = self.strategy.aggregate_evaluate(server_round, test_results, failures)
            val_metrics.update(test_metrics)
            if test_loss is not None:
                val_metrics[f"{MetricPrefix.TEST_PREFIX.value} loss - aggregated"] = test_loss
        return val_loss, val_metrics

    def _evaluate_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[float | None, dict[str, Scalar], EvaluateResultsAndFailures] | None:
        client_instructions = self.strategy.configure_evaluate(
            server_round=server_round, parameters=self.parameters, client_manager=self._client_manager
        )
        if not client_instructions:
            log(INFO, "evaluate_round %s: no clients selected, cancel", server_round)
            return None
        log(INFO, "evaluate_round %s: strategy sampled %s clients (out of %s)", server_round, len(client_instructions),
            self._client_manager.num_available())
        results, failures = self.transport.evaluate_clients(
            client_instructions, max_workers=self.max_workers, timeout=timeout, group_id=server_round
        )
        log(INFO, "evaluate_round %s received %s results and %s failures", server_round, len(results), len(failures))
        loss_aggregated, metrics_aggregated = self._handle_result_aggregation(server_round, results, failures)
        return loss_aggregated, metrics_aggregated, (results, failures)
