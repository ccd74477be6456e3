"""One-shot model-merge server (parity: ``fl4health/servers/model_merge_server.py:23-191``): one fit round to collect and
average client models, one federated + one centralized evaluation, optional latest-model checkpoint."""

from __future__ import annotations

import datetime
import timeit
from collections.abc import Sequence
from logging import INFO, WARNING
from typing import Any

from torch import nn

from fl4health_b200.checkpointing.checkpointer import LatestTorchModuleCheckpointer
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Parameters, Scalar, parameters_to_ndarrays
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.server import Server
from fl4health_b200.strategies.strategy import Strategy
from fl4health_b200.utils.random import generate_hash


class ModelMergeServer(Server):
    def __init__(
        self,
        client_manager: ClientManager,
        strategy: Strategy | None = None,
        checkpointer: LatestTorchModuleCheckpointer | None = None,
        server_model: nn.Module | None = None,
        parameter_exchanger: ParameterExchanger | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        server_name: str | None = None,
        transport: Any = None,
    ) -> None:
        assert (server_model is None and parameter_exchanger is None and checkpointer is None) or (
            server_model is not None and parameter_exchanger is not None and checkpointer is not None
        ), "checkpointer, server_model and parameter_exchanger must be given together or not at all"
        super().__init__(client_manager=client_manager, strategy=strategy, transport=transport)
        self.checkpointer, self.server_model, self.parameter_exchanger = checkpointer, server_model, parameter_exchanger
        self.server_name = server_name if server_name is not None else generate_hash()
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.server_name)

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        self.reports_manager.report({"fit_start": str(datetime.datetime.now()), "host_type": "server"})
        if num_rounds != 1:
            log(WARNING, "ModelMergeServer.fit performs exactly one merge round; num_rounds is ignored")
        history = History()
        self.parameters = Parameters([], "")
        log(INFO, "Federated Model Merging Starting")
        start = timeit.default_timer()
        res_fit = self.fit_round(server_round=1, timeout=timeout)
        if res_fit is not None:
            merged, fit_metrics, _ = res_fit
            if merged:
                self.parameters = merged
            history.add_metrics_distributed_fit(server_round=1, metrics=fit_metrics)
        else:
            log(WARNING, "Federated Model Merging Failed")
        res_fed = self.evaluate_round(server_round=1, timeout=timeout)
        if res_fed is not None and res_fed[1] is not None:
            history.add_metrics_distributed(server_round=1, metrics=res_fed[1])
        res_cen = self.strategy.evaluate(1, parameters=self.parameters)
        if res_cen is not None:
            history.add_metrics_centralized(server_round=1, metrics=res_cen[1])
        self._maybe_checkpoint(loss_aggregated=0.0, metrics_aggregated={}, server_round=1)
        self.reports_manager.report({"fit_end": str(datetime.datetime.now()), "metrics_centralized": history.metrics_centralized,
                                     "losses_centralized": history.losses_centralized, "host_type": "server"})
        elapsed = timeit.default_timer() - start
        log(INFO, "Federated Model Merging Finished in %s", elapsed)
        return history, elapsed

    def _hydrate_model_for_checkpointing(self) -> nn.Module:
        assert self.server_model is not None and self.parameter_exchanger is not None
        self.parameter_exchanger.pull_parameters(parameters_to_ndarrays(self.parameters), self.server_model)
        return self.server_model

    def _maybe_checkpoint(self, loss_aggregated: float, metrics_aggregated: dict[str, Scalar], server_round: int) -> None:
        if self.checkpointer and self.server_model and self.parameter_exchanger:
            self.checkpointer.maybe_checkpoint(self._hydrate_model_for_checkpointing(), loss_aggregated, metrics_aggregated)
        else:
            log(WARNING, "Server model is not being checkpointed: checkpointer, server_model and parameter_exchanger are all required")

    def shutdown(self) -> None:
        self.reports_manager.shutdown()
