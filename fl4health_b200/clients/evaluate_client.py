"""Evaluation-only client (parity: ``fl4health/clients/evaluate_client.py:24-282``): evaluates a locally stored
checkpoint and/or the global model sent by the server; ``fit`` / ``get_parameters`` are errors."""

from __future__ import annotations

import datetime
from collections.abc import Sequence
from logging import INFO, WARNING
from pathlib import Path

import torch
from torch import nn
from torch.nn.modules.loss import _Loss
from torch.utils.data import DataLoader

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.metrics.metric_managers import MetricManager
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.utils.client import move_data_to_device
from fl4health_b200.utils.losses import EvaluationLosses, LossMeter, LossMeterType
from fl4health_b200.utils.random import generate_hash


class EvaluateClient(BasicClient):
    def __init__(
        self,
        data_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        model_checkpoint_path: Path | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        client_name: str | None = None,
    ) -> None:
        # deliberately NOT calling BasicClient.__init__: there is no training machinery in this client
        self.client_name = generate_hash() if client_name is None else client_name
        self.data_path = data_path
        self.device = torch.device(device)
        self.model_checkpoint_path = model_checkpoint_path
        self.metrics = metrics
        self.initialized = False
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.client_name)
        self.global_loss_meter = LossMeter[EvaluationLosses](loss_meter_type, EvaluationLosses)
        self.global_metric_manager = MetricManager(self.metrics, "global_eval_manager")
        self.local_loss_meter = LossMeter[EvaluationLosses](loss_meter_type, EvaluationLosses)
        self.local_metric_manager = MetricManager(self.metrics, "local_eval_manager")
        self.data_loader: DataLoader
        self.criterion: _Loss
        self.local_model: nn.Module | None = None
        self.global_model: nn.Module | None = None

    def get_parameters(self, config: Config) -> NDArrays:
        raise ValueError("Get Parameters is not implemented for an Evaluation-Only Client")

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        raise ValueError("Fit is not implemented for an Evaluation-Only Client")

    def setup_client(self, config: Config) -> None:
        (self.data_loader,) = self.get_data_loader(config)
        self.global_model = self.initialize_global_model(config)
        self.local_model = self.get_local_model(config)
        self.num_samples = len(self.data_loader.dataset)  # type: ignore[arg-type]
        self.criterion = self.get_criterion(config)
        self.parameter_exchanger = self.get_parameter_exchanger(config)
        self.reports_manager.report({"host_type": "client", "initialized": str(datetime.datetime.now())})
        self.initialized = True

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        assert not fitting_round
        if len(parameters) > 0:
            assert self.global_model is not None and self.parameter_exchanger is not None
            self.parameter_exchanger.pull_parameters(parameters, self.global_model, config)
        else:
            self.global_model = None  # the server sent nothing: only the local checkpoint is evaluated

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        start = datetime.datetime.now()
        self.set_parameters(parameters, config, fitting_round=False)
        assert self.local_model or self.global_model
        loss, metric_values = self.validate()
        end = datetime.datetime.now()
        self.reports_manager.report(
            {"eval_metrics": metric_values, "eval_loss": loss, "eval_start": str(start),
             "eval_time_elapsed": str(end - start), "eval_end": str(end)}, 0)
        return loss, self.num_samples, metric_values

    def _handle_logging(self, losses: EvaluationLosses, metrics_dict: dict[str, Scalar], is_global: bool) -> None:  # type: ignore[override]
        prefix = "Global Model" if is_global else "Local Model"
        log(INFO, f"Client Evaluation {prefix} Losses: {losses.as_dict()} | Metrics: {metrics_dict}")

    def validate_on_model(self, model: nn.Module, metric_meter: MetricManager, loss_meter: LossMeter,
                          is_global: bool) -> tuple[EvaluationLosses, dict[str, Scalar]]:
        model.eval()
        metric_meter.clear()
        loss_meter.clear()
        model.to(self.device)
        with torch.no_grad():
            for inputs, targets in self.data_loader:
                inputs, targets = move_data_to_device(inputs, self.device), move_data_to_device(targets, self.device)
                preds = {"prediction": model(inputs)}
                losses = self.compute_evaluation_loss(preds, {}, targets)
                metric_meter.update(preds, targets)
                loss_meter.update(losses)
        metrics, losses = metric_meter.compute(), loss_meter.compute()
        self._handle_logging(losses, metrics, is_global)
        return losses, metrics

    def validate(self, include_loss_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        local_loss = local_metrics = global_loss = global_metrics = None
        if self.local_model:
            log(INFO, "Performing evaluation on local model")
            local_loss, local_metrics = self.validate_on_model(self.local_model, self.local_metric_manager, self.local_loss_meter, False)
        if self.global_model:
            log(INFO, "Performing evaluation on global model")
            global_loss, global_metrics = self.validate_on_model(self.global_model, self.global_metric_manager, self.global_loss_meter, True)
        metrics = EvaluateClient.merge_metrics(global_metrics, local_metrics)
        if global_loss:
            metrics.update({f"global_loss_{k}": v for k, v in global_loss.as_dict().items()})
        if local_loss:
            metrics.update({f"local_loss_{k}": v for k, v in local_loss.as_dict().items()})
        return float("nan"), metrics  # no single loss is meaningful across local/global models

    @staticmethod
    def merge_metrics(global_metrics: dict[str, Scalar] | None, local_metrics: dict[str, Scalar] | None) -> dict[str, Scalar]:
        if global_metrics:
            metrics = global_metrics
            for name, value in (local_metrics or {}).items():
                if name in metrics:
                    log(WARNING, f"metric_name: {name} already exists in dictionary. Please ensure that this is intended behavior")
                metrics[name] = value
            return metrics
        if local_metrics:
            return local_metrics
        raise ValueError("Both metric dictionaries are None. At least one global or local model should be present.")

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchanger()

    def get_data_loader(self, config: Config) -> tuple[DataLoader]:
        raise NotImplementedError

    def initialize_global_model(self, config: Config) -> nn.Module | None:
        return None

    def get_local_model(self, config: Config) -> nn.Module | None:
        if self.model_checkpoint_path:
            log(INFO, f"Loading model checkpoint at: {self.model_checkpoint_path}")
            return torch.load(self.model_checkpoint_path, weights_only=False)
        return None
