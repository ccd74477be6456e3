"""FedRep (Collins et al. 2021): per round, train the head with the representation frozen, then the representation
with the head frozen; only the representation is exchanged (parity: ``fl4health/clients/fedrep_client.py:33-429``).
Config keys: ``local_head_steps`` + ``local_rep_steps`` or ``local_head_epochs`` + ``local_rep_epochs``."""

from __future__ import annotations

import datetime
from collections.abc import Sequence
from enum import Enum
from logging import INFO
from pathlib import Path

import torch
from torch.optim import Optimizer

from fl4health_b200.checkpointing.client_module import CheckpointMode, ClientCheckpointAndStateModule
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.engine.options import EngineOptions
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.fedrep_base import FedRepModel
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.losses import LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchInputType, TorchPredType, TorchTargetType

EpochsAndStepsTuple = tuple[int | None, int | None, int | None, int | None]


class FedRepTrainMode(Enum):
    HEAD = "head"
    REPRESENTATION = "representation"


class FedRepClient(BasicClient):
    def __init__(
        self,
        data_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        checkpoint_and_state_module: ClientCheckpointAndStateModule | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        progress_bar: bool = False,
        client_name: str | None = None,
        engine_options: EngineOptions | None = None,
    ) -> None:
        super().__init__(
            data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
            checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters, progress_bar=progress_bar,
            client_name=client_name, engine_options=engine_options,
        )
        self.fedrep_train_mode = FedRepTrainMode.HEAD

    def _graph_variant(self) -> object:
        return self.fedrep_train_mode.value

    def _prepare_train_representations(self) -> None:
        assert isinstance(self.model, FedRepModel)
        self.fedrep_train_mode = FedRepTrainMode.REPRESENTATION
        self.model.unfreeze_base_module()
        self.model.freeze_head_module()

    def _prepare_train_head(self) -> None:
        assert isinstance(self.model, FedRepModel)
        self.fedrep_train_mode = FedRepTrainMode.HEAD
        self.model.unfreeze_head_module()
        self.model.freeze_base_module()

    def _prefix_loss_and_metrics_dictionaries(self, prefix: str, loss_dict: dict[str, float], metrics_dict: dict[str, Scalar]) -> None:
        for key in list(loss_dict):
            loss_dict[f"{prefix}_{key}"] = loss_dict.pop(key)
        for key in list(metrics_dict):
            metrics_dict[f"{prefix}_{key}"] = metrics_dict.pop(key)

    def _extract_epochs_or_steps_specified(self, config: Config) -> EpochsAndStepsTuple:
        epochs = ("local_head_epochs" in config) and ("local_rep_epochs" in config)
        steps = ("local_head_steps" in config) and ("local_rep_steps" in config)
        if epochs and steps:
            raise ValueError("Cannot specify both epochs and steps based training values in the config")
        if epochs:
            return narrow_dict_type(config, "local_head_epochs", int), narrow_dict_type(config, "local_rep_epochs", int), None, None
        if steps:
            return None, None, narrow_dict_type(config, "local_head_steps", int), narrow_dict_type(config, "local_rep_steps", int)
        raise ValueError(
            "Keys should be one of {local_head_epochs, local_rep_epochs} or {local_head_steps, local_rep_steps}"
        )

    def process_fed_rep_config(self, config: Config) -> tuple[EpochsAndStepsTuple, int, bool]:
        current_server_round = narrow_dict_type(config, "current_server_round", int)
        return self._extract_epochs_or_steps_specified(config), current_server_round, bool(config.get("evaluate_after_fit", False))

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError('Return a dict with keys "representation" and "head"')

    def set_optimizer(self, config: Config) -> None:
        optimizers = self.get_optimizer(config)
        assert isinstance(optimizers, dict) and set(optimizers.keys()) == {"representation", "head"}, (
            'Optimizer keys must be "representation" and "head" to use FedRep'
        )
        self.optimizers = optimizers

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, SequentiallySplitExchangeBaseModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())

    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        round_start = datetime.datetime.now()
        (head_epochs, rep_epochs, head_steps, rep_steps), current_server_round, evaluate_after_fit = self.process_fed_rep_config(config)
        if not self.initialized:
            self.setup_client(config)
        self.set_parameters(parameters, config, fitting_round=True)
        self.update_before_train(current_server_round)
        fit_start = datetime.datetime.now()
        if head_epochs and rep_epochs:
            loss_dict, metrics = self.train_fedrep_by_epochs(head_epochs, rep_epochs, current_server_round)
        elif head_steps and rep_steps:
            loss_dict, metrics = self.train_fedrep_by_steps(head_steps, rep_steps, current_server_round)
        else:
            raise ValueError(f"Local epochs or steps not correctly specified: {head_epochs}, {rep_epochs}, {head_steps}, {rep_steps}")
        fit_time = datetime.datetime.now() - fit_start
        if self._should_evaluate_after_fit(evaluate_after_fit):
            validation_loss, validation_metrics = self.validate()
            metrics.update(validation_metrics)
            self._maybe_checkpoint(validation_loss, validation_metrics, CheckpointMode.PRE_AGGREGATION)
        self.reports_manager.report(
            {"fit_metrics": metrics, "fit_losses": loss_dict, "round": current_server_round,
             "round_start": str(round_start), "fit_time_elapsed": str(fit_time)},
            current_server_round,
        )
        return self.get_parameters(config), self.num_train_samples, metrics

    def _two_phase(self, run_phase, head_amount: int, rep_amount: int, unit: str, current_round: int | None):  # noqa: ANN001, ANN202
        self._prepare_train_head()
        log(INFO, f"Beginning FedRep Head Training Phase for {head_amount} {unit}")
        loss_head, metrics_head = run_phase(head_amount, current_round)
        self._prefix_loss_and_metrics_dictionaries("head", loss_head, metrics_head)
        self._prepare_train_representations()
        log(INFO, f"Beginning FedRep Representation Training Phase for {rep_amount} {unit}")
        loss_rep, metrics_rep = run_phase(rep_amount, current_round)
        self._prefix_loss_and_metrics_dictionaries("rep", loss_rep, metrics_rep)
        loss_head.update(loss_rep)
        metrics_head.update(metrics_rep)
        return loss_head, metrics_head

    def train_fedrep_by_epochs(self, head_epochs: int, rep_epochs: int, current_round: int | None = None) -> tuple[dict[str, float], dict[str, Scalar]]:
        return self._two_phase(self.train_by_epochs, head_epochs, rep_epochs, "Epochs", current_round)

    def train_fedrep_by_steps(self, head_steps: int, rep_steps: int, current_round: int | None = None) -> tuple[dict[str, float], dict[str, Scalar]]:
        assert isinstance(self.model, FedRepModel)
        return self._two_phase(self.train_by_steps, head_steps, rep_steps, "Steps", current_round)

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        self.optimizers["representation"].zero_grad()
        self.optimizers["head"].zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            target = self.transform_target(target)
            losses = self.compute_training_loss(preds, features, target)
        losses.backward["backward"].backward()
        if self.fedrep_train_mode == FedRepTrainMode.HEAD:
            self.optimizers["head"].step()
        elif self.fedrep_train_mode == FedRepTrainMode.REPRESENTATION:
            self.optimizers["representation"].step()
        else:
            raise ValueError("Training Mode in an invalid state")
        return losses, preds
