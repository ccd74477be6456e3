"""``FlServer``: the federated round loop with reporting, checkpointing, resume and a failure policy.

API parity: ``fl4health/servers/base_server.py:36-643`` (constructor, ``fit``, ``fit_round``, ``evaluate_round``,
``poll_clients_for_sample_counts``, ``update_before_fit``, ``shutdown``, test/val metric split, report keys).
The round loop is an explicit phase machine (``servers/round_machine.py``) shared by the in-process simulation and the
SPMD runtimes; ``FlServer`` supplies the content of each phase.  It is driven directly (no Flower, no gRPC): clients are reached through ``self.transport`` which is either the
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
from fl4health_b200.engine.round_protocol import Stopwatch
from fl4health_b200.metrics.base_metrics import TEST_LOSS_KEY, TEST_NUM_EXAMPLES_KEY, MetricPrefix
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.servers.round_machine import RoundMachine, inject_fault_after
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
        self.fl_config, self.accept_failures = fl_config, accept_failures
        self.on_init_parameters_config_fn = on_init_parameters_config_fn
        self.checkpoint_and_state_module = checkpoint_and_state_module or BaseServerCheckpointAndStateModule(
            model=None, parameter_exchanger=None, model_checkpointers=None, state_checkpointer=None
        )
        self.server_name = server_name or generate_hash()
        log(INFO, f"Server Name: {self.server_name}")
        self.current_round: int
        self.history: History
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.server_name)
        # observers called with the round number once a round (fit + evaluate) is complete
        self.round_end_hooks: list[Callable[[int], None]] = []
        self._log_fl_config()

    # ------------------------------------------------------------------------------------------------------
    # the run: phase machine (servers/round_machine.py) + run-level reports
    # ------------------------------------------------------------------------------------------------------
    def update_before_fit(self, num_rounds: int, timeout: float | None) -> None:
        """Hook run once before the round loop (plan negotiation, warm starts...)."""

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        clock = Stopwatch()
        self.reports_manager.report({"fit_start": str(clock.mark("begin")), "host_type": "server"})
        self.update_before_fit(num_rounds, timeout)
        resumable = self.checkpoint_and_state_module.state_checkpointer is not None
        history, seconds = self._run_rounds(num_rounds, timeout, resumable=resumable)
        clock.mark("end")
        span = clock.span("fit", "begin", "end")
        self.reports_manager.report({"fit_elapsed_time": span["fit_time_elapsed"], "fit_end": span["fit_end"],
                                     "num_rounds": num_rounds, "host_type": "server"})
        self.report_centralized_eval(history, num_rounds)
        log(INFO, "[SUMMARY]\n%s", history)
        return history, seconds

    def _run_rounds(self, num_rounds: int, timeout: float | None, resumable: bool) -> tuple[History, float]:
        return RoundMachine(self, num_rounds, timeout, resumable).run()

    def fit_with_per_round_checkpointing(self, num_rounds: int, timeout: float | None) -> tuple[History, float]:
        """Resume-able loop: reloads ``current_round``/``history``/parameters if a state file exists and saves the
        server state after every round."""
        return self._run_rounds(num_rounds, timeout, resumable=True)

    def report_centralized_eval(self, history: History, num_rounds: int) -> None:
        """Replay the centrally evaluated losses / metrics of rounds 1..num_rounds to the reporters."""
        losses = {rnd: loss for rnd, loss in history.losses_centralized if 1 <= rnd <= num_rounds}
        per_metric = {name: dict(points) for name, points in history.metrics_centralized.items()}
        for server_round in sorted(losses):
            self.reports_manager.report({"val - loss - centralized": losses[server_round]}, server_round)
            snapshot = {name: points[server_round] for name, points in per_metric.items() if server_round in points}
            self.reports_manager.report({"eval_round_metrics_centralized": snapshot}, server_round)

    def _maybe_inject_fault(self) -> None:
        inject_fault_after(self.current_round - 1)

    def shutdown(self) -> None:
        self.reports_manager.report({"shutdown": str(datetime.datetime.now())})
        self.reports_manager.shutdown()

    # ------------------------------------------------------------------------------------------------------
    # round phases
    # ------------------------------------------------------------------------------------------------------
    def _enforce_failure_policy(self, failures: FitFailures | EvaluateFailures | None, timeout: float | None) -> None:
        if failures and not self.accept_failures:
            self._log_client_failures(failures)
            self._terminate_after_unacceptable_failures(timeout)

    def fit_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[Parameters | None, dict[str, Scalar], FitResultsAndFailures] | None:
        clock = Stopwatch()
        clock.mark("begin")
        outcome = super().fit_round(server_round, timeout)
        clock.mark("end")
        timing = clock.span("fit_round", "begin", "end")
        timing["fit_round_seconds"] = (clock.marks["end"] - clock.marks["begin"]).total_seconds()
        self.reports_manager.report(timing, server_round)
        if outcome is not None:
            _, metrics, results_and_failures = outcome
            self.reports_manager.report({"fit_round_metrics": metrics}, server_round)
            self._enforce_failure_policy(results_and_failures[1] if results_and_failures else None, timeout)
        return outcome

    def evaluate_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[float | None, dict[str, Scalar], EvaluateResultsAndFailures] | None:
        clock = Stopwatch()
        clock.mark("begin")
        outcome = self._evaluate_round(server_round, timeout)
        clock.mark("end")
        if not outcome:
            return outcome
        loss, metrics, (_, failures) = outcome
        self._enforce_failure_policy(failures, timeout)
        if loss is None:
            return outcome
        self._maybe_checkpoint(loss, metrics, server_round)
        payload: dict[str, Any] = {"val - loss - aggregated": loss, "round": server_round,
                                   **clock.span("eval_round", "begin", "end")}
        for unit, key in (("fit_epoch", "local_epochs"), ("fit_step", "local_steps")):
            if self.fl_config.get(key) is not None:  # progress in the clients' own unit (epochs win over steps)
                payload[unit] = server_round * self.fl_config[key]  # type: ignore[operator]
                break
        self.reports_manager.report(payload, server_round)
        if metrics:
            self.reports_manager.report({"eval_round_metrics_aggregated": metrics}, server_round)
        return outcome

    def _evaluate_round(
        self, server_round: int, timeout: float | None
    ) -> tuple[float | None, dict[str, Scalar], EvaluateResultsAndFailures] | None:
        instructions = self.strategy.configure_evaluate(
            server_round=server_round, parameters=self.parameters, client_manager=self._client_manager
        )
        if not instructions:
            log(INFO, "evaluate_round %s: no clients selected, cancel", server_round)
            return None
        log(INFO, "evaluate_round %s: strategy sampled %s clients (out of %s)", server_round, len(instructions),
            self._client_manager.num_available())
        results, failures = self.transport.evaluate_clients(
            instructions, max_workers=self.max_workers, timeout=timeout, group_id=server_round
        )
        log(INFO, "evaluate_round %s received %s results and %s failures", server_round, len(results), len(failures))
        loss, metrics = self._handle_result_aggregation(server_round, results, failures)
        return loss, metrics, (results, failures)

    def _unpack_metrics(
        self, results: list[tuple[ClientProxy, EvaluateRes]]
    ) -> tuple[list[tuple[ClientProxy, EvaluateRes]], list[tuple[ClientProxy, EvaluateRes]]]:
        """Split each client's metrics into validation and ``"test - "``-prefixed parts (own loss / sample count)."""
        prefix = MetricPrefix.TEST_PREFIX.value
        validation, testing = [], []
        for proxy, res in results:
            split: dict[bool, dict[str, Scalar]] = {True: {}, False: {}}
            for name, value in res.metrics.items():
                split[name.startswith(prefix)][name] = value
            validation.append((proxy, EvaluateRes(res.status, res.loss, res.num_examples, split[False])))
            held_out = split[True]
            if not held_out:
                continue
            assert TEST_LOSS_KEY in held_out and TEST_NUM_EXAMPLES_KEY in held_out, (
                f"'{TEST_NUM_EXAMPLES_KEY}' and '{TEST_LOSS_KEY}' keys must be present in test_metrics for aggregation"
            )
            loss = float(held_out.pop(TEST_LOSS_KEY))  # type: ignore[arg-type]
            count = int(held_out.pop(TEST_NUM_EXAMPLES_KEY))  # type: ignore[arg-type]
            testing.append((proxy, EvaluateRes(res.status, loss, count, held_out)))
        return validation, testing

    def _handle_result_aggregation(
        self, server_round: int, results: list[tuple[ClientProxy, EvaluateRes]], failures: list[Any]
    ) -> tuple[float | None, dict[str, Scalar]]:
        validation, testing = self._unpack_metrics(results)
        loss, metrics = self.strategy.aggregate_evaluate(server_round, validation, failures)
        if testing:
            test_loss, test_metrics = self.strategy.aggregate_evaluate(server_round, testing, failures)
            metrics.update(test_metrics)
            if test_loss is not None:
                metrics[f"{MetricPrefix.TEST_PREFIX.value} loss - aggregated"] = test_loss
        return loss, metrics

    def poll_clients_for_sample_counts(self, timeout: float | None) -> list[int]:
        log(INFO, "Polling Clients for sample counts")
        assert isinstance(self.strategy, StrategyWithPolling)
        instructions = self.strategy.configure_poll(server_round=1, client_manager=self._client_manager)
        answers, _ = self.transport.poll_clients(instructions, max_workers=self.max_workers, timeout=timeout)
        counts = [int(res.properties["num_train_samples"]) for _, res in answers]
        log(INFO, f"Polling complete: Retrieved {len(counts)} sample counts")
        return counts

    # ------------------------------------------------------------------------------------------------------
    # initial parameters
    # ------------------------------------------------------------------------------------------------------
    def _get_initial_parameters(self, server_round: int, timeout: float | None) -> Parameters:
        from_strategy = self.strategy.initialize_parameters(client_manager=self._client_manager)
        if from_strategy is not None:
            log(INFO, "Using initial global parameters provided by strategy")
            return from_strategy
        log(INFO, "Requesting initial parameters from one random client")
        manager = self._client_manager
        donor = (manager.sample_one() if isinstance(manager, BaseFractionSamplingManager) else manager.sample(1))[0]
        config: dict[str, Scalar] = {}
        if self.on_init_parameters_config_fn is not None:
            config = self.on_init_parameters_config_fn(server_round)
        else:
            log(WARNING, "on_init_parameters_config_fn is None; clients usually need a config to set themselves up.")
        res = self.transport_get_parameters(donor, GetParametersIns(config=config), timeout, server_round)
        if res.status.code == Code.OK:
            log(INFO, "Received initial parameters from one random client")
        else:
            log(WARNING, "Failed to receive initial parameters from the client. Empty initial parameters will be used.")
        # own copy: the global model must not alias the donor's live training buffers (its arena)
        owned = [t.detach().clone() if isinstance(t, torch.Tensor) else t for t in res.parameters.tensors]
        initial = Parameters(tensors=owned, tensor_type=res.parameters.tensor_type)
        if isinstance(self.strategy, BasicFedAvg):
            self.strategy.add_auxiliary_information(initial)
        return initial

    def transport_get_parameters(self, proxy: ClientProxy, ins: GetParametersIns, timeout: float | None, server_round: int) -> Any:
        via_transport = getattr(self.transport, "get_parameters", None)
        if via_transport is not None:
            return via_transport(proxy, ins, timeout, server_round)
        return proxy.get_parameters(ins=ins, timeout=timeout, group_id=server_round)

    # ------------------------------------------------------------------------------------------------------
    # persistence, failure handling, logging
    # ------------------------------------------------------------------------------------------------------
    def _save_server_state(self) -> None:
        assert self.checkpoint_and_state_module.state_checkpointer is not None
        if self.transport.is_coordinator():
            self.checkpoint_and_state_module.save_state(self, self.parameters)

    def _load_server_state(self) -> bool:
        assert self.checkpoint_and_state_module.state_checkpointer is not None
        restored = self.checkpoint_and_state_module.maybe_load_state(self)
        if not restored:
            return False
        self.parameters = restored
        log(INFO, "Loaded server state from checkpoint")
        return True

    def _maybe_checkpoint(self, loss_aggregated: float, metrics_aggregated: dict[str, Scalar], server_round: int) -> None:
        if self.transport.is_coordinator():
            self.checkpoint_and_state_module.maybe_checkpoint(self.parameters, loss_aggregated, metrics_aggregated)

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
                continue
            log(ERROR, f"Client {failure[0].cid} failed but did not return an exception. Partial results were received")

    def _log_fl_config(self) -> None:
        log(INFO, "FL Configuration:" if self.fl_config else "FL Config is Empty")
        for key, value in self.fl_config.items():
            if not isinstance(value, bytes):
                log(INFO, f"Key: {key} Value: {value!r}")
