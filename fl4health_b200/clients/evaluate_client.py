"""Evaluation-only client (parity: ``fl4health/clients/evaluate_client.py:24-282``): evaluates a locally stored
checkpoint and/or the global model sent by the server; ``fit`` / ``get_parameters`` are errors."""

from __future__ import annotations

import datetime
from collections.abc import Sequence
from dataclasses import dataclass
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


@dataclass
class _Subject:
    """One network under evaluation ("local" checkpoint / "global" server model) with its own meters."""

    label: str
    loss_meter: LossMeter
    metric_manager: MetricManager
    model: nn.Module | None = None

    @property
    def title(self) -> str:
        return f"{self.label.capitalize()} Model"


class EvaluateClient(BasicClient):
    """No training machinery at all (``BasicClient.__init__`` is deliberately not called): the client owns a table of
    evaluation *subjects* and scores whichever of them exist on one data loader."""

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
        self.client_name = client_name if client_name is not None else generate_hash()
        self.data_path, self.metrics, self.model_checkpoint_path = data_path, metrics, model_checkpoint_path
        self.device = torch.device(device)
        self.initialized = False
        self.reports_manager = ReportsManager(reporters)
        self.reports_manager.initialize(id=self.client_name)
        self._subjects = {
            label: _Subject(label, LossMeter[EvaluationLosses](loss_meter_type, EvaluationLosses),
                            MetricManager(self.metrics, f"{label}_eval_manager"))
            for label in ("local", "global")
        }
        self.data_loader: DataLoader
        self.criterion: _Loss

    # the reference's attribute names, backed by the subject table
    local_model = property(lambda self: self._subjects["local"].model,
                           lambda self, model: setattr(self._subjects["local"], "model", model))
    global_model = property(lambda self: self._subjects["global"].model,
                            lambda self, model: setattr(self._subjects["global"], "model", model))
    local_loss_meter = property(lambda self: self._subjects["local"].loss_meter)
    global_loss_meter = property(lambda self: self._subjects["global"].loss_meter)
    local_metric_manager = property(lambda self: self._subjects["local"].metric_manager)
    global_metric_manager = property(lambda self: self._subjects["global"].metric_manager)

    # ---------------------------------------------------------------------------------------------- protocol
    def get_parameters(self, config: Config) -> NDArrays:
        raise ValueError("Get Parameters is not implemented for an Evaluation-Only Client")

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        raise ValueError("Fit is not implemented for an Evaluation-Only Client")

    def setup_client(self, config: Config) -> None:
        (self.data_loader,) = self.get_data_loader(config)
        self.num_samples = len(self.data_loader.dataset)  # type: ignore[arg-type]
        self.global_model, self.local_model = self.initialize_global_model(config), self.get_local_model(config)
        self.criterion = self.get_criterion(config)
        self.parameter_exchanger = self.get_parameter_exchanger(config)
        self.reports_manager.report({"host_type": "client", "initialized": str(datetime.datetime.now())})
        self.initialized = True

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        assert not fitting_round
        if not len(parameters):
            self.global_model = None  # the server sent nothing: only the local checkpoint is evaluated
            return
        assert self.global_model is not None and self.parameter_exchanger is not None
        self.parameter_exchanger.pull_parameters(parameters, self.global_model, config)

    def evaluate(self, parameters: NDArrays, config: Config) -> tuple[float, int, dict[str, Scalar]]:
        if not self.initialized:
            self.setup_client(config)
        began = datetime.datetime.now()
        self.set_parameters(parameters, config, fitting_round=False)
        assert self.local_model or self.global_model
        loss, scores = self.validate()
        ended = datetime.datetime.now()
        self.reports_manager.report({"eval_metrics": scores, "eval_loss": loss, "eval_start": str(began),
                                     "eval_time_elapsed": str(ended - began), "eval_end": str(ended)}, 0)
        return loss, self.num_samples, scores

    # ---------------------------------------------------------------------------------------------- scoring
    def _handle_logging(self, losses: EvaluationLosses, metrics_dict: dict[str, Scalar], is_global: bool) -> None:  # type: ignore[override]
        title = self._subjects["global" if is_global else "local"].title
        log(INFO, f"Client Evaluation {title} Losses: {losses.as_dict()} | Metrics: {metrics_dict}")

    def validate_on_model(self, model: nn.Module, metric_meter: MetricManager, loss_meter: LossMeter,
                          is_global: bool) -> tuple[EvaluationLosses, dict[str, Scalar]]:
        model.to(self.device).eval()
        metric_meter.clear()
        loss_meter.clear()
        with torch.no_grad():
            for batch_input, batch_target in self.data_loader:
                batch_input = move_data_to_device(batch_input, self.device)
                batch_target = move_data_to_device(batch_target, self.device)
                preds = {"prediction": model(batch_input)}
                loss_meter.update(self.compute_evaluation_loss(preds, {}, batch_target))
                metric_meter.update(preds, batch_target)
        scores, losses = metric_meter.compute(), loss_meter.compute()
        self._handle_logging(losses, scores, is_global)
        return losses, scores

    def validate(self, include_loss_in_metrics: bool = False) -> tuple[float, dict[str, Scalar]]:
        outcome: dict[str, tuple[EvaluationLosses, dict[str, Scalar]]] = {}
        for label in ("local", "global"):
            subject = self._subjects[label]
            if subject.model:
                log(INFO, f"Performing evaluation on {label} model")
                outcome[label] = self.validate_on_model(subject.model, subject.metric_manager, subject.loss_meter,
                                                        is_global=label == "global")
        merged = EvaluateClient.merge_metrics(outcome.get("global", (None, None))[1], outcome.get("local", (None, None))[1])
        for label in ("global", "local"):
            if label in outcome and outcome[label][0]:
                merged.update({f"{label}_loss_{name}": value for name, value in outcome[label][0].as_dict().items()})
        return float("nan"), merged  # no single loss is meaningful across local/global models

    @staticmethod
    def merge_metrics(global_metrics: dict[str, Scalar] | None, local_metrics: dict[str, Scalar] | None) -> dict[str, Scalar]:
        if not global_metrics and not local_metrics:
            raise ValueError("Both metric dictionaries are None. At least one global or local model should be present.")
        if not global_metrics:
            return local_metrics  # type: ignore[return-value]
        for name in set(global_metrics) & set(local_metrics or {}):
            log(WARNING, f"metric_name: {name} already exists in dictionary. Please ensure that this is intended behavior")
        global_metrics.update(local_metrics or {})
        return global_metrics

    # ---------------------------------------------------------------------------------------------- factories
    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchanger()

    def get_data_loader(self, config: Config) -> tuple[DataLoader]:
        raise NotImplementedError

    def initialize_global_model(self, config: Config) -> nn.Module | None:
        return None

    def get_local_model(self, config: Config) -> nn.Module | None:
        if not self.model_checkpoint_path:
            return None
        log(INFO, f"Loading model checkpoint at: {self.model_checkpoint_path}")
        return torch.load(self.model_checkpoint_path, weights_only=False)
