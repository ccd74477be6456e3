"""FedRep (Collins et al. 2021): per round, train the head with the representation frozen, then the representation
with the head frozen; only the representation is exchanged (parity: ``fl4health/clients/fedrep_client.py:33-429``).
Config keys: ``local_head_steps`` + ``local_rep_steps`` or ``local_head_epochs`` + ``local_rep_epochs``."""

from __future__ import annotations

import datetime
from dataclasses import dataclass
from enum import Enum
from logging import INFO
from typing import Any

from torch.optim import Optimizer

from fl4health_b200.checkpointing.client_module import CheckpointMode
from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, Scalar
from fl4health_b200.model_bases.fedrep_base import FedRepModel
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.losses import TrainingLosses
from fl4health_b200.utils.typing import TorchInputType, TorchPredType, TorchTargetType

EpochsAndStepsTuple = tuple[int | None, int | None, int | None, int | None]


class FedRepTrainMode(Enum):
    HEAD = "head"
    REPRESENTATION = "representation"


# phase -> (optimizer that steps, log label, prefix of the reported losses / metrics)
_PHASES = {
    FedRepTrainMode.HEAD: ("head", "Head", "head"),
    FedRepTrainMode.REPRESENTATION: ("representation", "Representation", "rep"),
}


@dataclass(frozen=True)
class _PhaseBudget:
    """How much local work each of the two phases gets, and in which unit."""

    unit: str  # "epochs" | "steps"
    head: int
    representation: int

    @classmethod
    def from_config(cls, config: Config) -> "_PhaseBudget":
        offered = [unit for unit in ("epochs", "steps")
                   if f"local_head_{unit}" in config and f"local_rep_{unit}" in config]
        if len(offered) == 2:
            raise ValueError("Cannot specify both epochs and steps based training values in the config")
        if not offered:
            raise ValueError("Keys should be one of {local_head_epochs, local_rep_epochs} or {local_head_steps, local_rep_steps}")
        unit = offered[0]
        return cls(unit, narrow_dict_type(config, f"local_head_{unit}", int), narrow_dict_type(config, f"local_rep_{unit}", int))

    def legacy_tuple(self) -> EpochsAndStepsTuple:
        pair = (self.head, self.representation)
        return (*pair, None, None) if self.unit == "epochs" else (None, None, *pair)


class FedRepClient(BasicClient):
    """Constructor arguments are ``BasicClient``'s.  ``get_optimizer`` must return ``{"representation": ..., "head": ...}``."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.fedrep_train_mode = FedRepTrainMode.HEAD

    def _graph_variant(self) -> object:
        return self.fedrep_train_mode.value  # the two phases launch different optimizer kernels: one graph each

    # ------------------------------------------------------------------------------------------ wiring
    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError('Return a dict with keys "representation" and "head"')

    def set_optimizer(self, config: Config) -> None:
        per_part = self.get_optimizer(config)
        expected = {keys[0] for keys in _PHASES.values()}
        assert isinstance(per_part, dict) and set(per_part) == expected, 'Optimizer keys must be "representation" and "head" to use FedRep'
        self.optimizers = per_part

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, SequentiallySplitExchangeBaseModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())

    # ------------------------------------------------------------------------------------------ phases
    def _enter_phase(self, mode: FedRepTrainMode) -> None:
        """Freeze the half of the network that is not being trained in ``mode``."""
        assert isinstance(self.model, FedRepModel)
        self.fedrep_train_mode = mode
        if mode is FedRepTrainMode.HEAD:
            self.model.unfreeze_head_module()
            self.model.freeze_base_module()
        else:
            self.model.unfreeze_base_module()
            self.model.freeze_head_module()

    def _prepare_train_head(self) -> None:
        self._enter_phase(FedRepTrainMode.HEAD)

    def _prepare_train_representations(self) -> None:
        self._enter_phase(FedRepTrainMode.REPRESENTATION)

    def _prefix_loss_and_metrics_dictionaries(self, prefix: str, loss_dict: dict[str, float], metrics_dict: dict[str, Scalar]) -> None:
        for table in (loss_dict, metrics_dict):
            for key in list(table):
                table[f"{prefix}_{key}"] = table.pop(key)

    def _extract_epochs_or_steps_specified(self, config: Config) -> EpochsAndStepsTuple:
        return _PhaseBudget.from_config(config).legacy_tuple()

    def process_fed_rep_config(self, config: Config) -> tuple[EpochsAndStepsTuple, int, bool]:
        server_round = narrow_dict_type(config, "current_server_round", int)
        return self._extract_epochs_or_steps_specified(config), server_round, bool(config.get("evaluate_after_fit", False))

    def _run_phases(self, budget: _PhaseBudget, current_round: int | None) -> tuple[dict[str, float], dict[str, Scalar]]:
        """Head phase then representation phase, each through the ordinary epoch / step driver; results are merged under
        ``head_`` / ``rep_`` prefixes."""
        driver = self.train_by_epochs if budget.unit == "epochs" else self.train_by_steps
        losses: dict[str, float] = {}
        metrics: dict[str, Scalar] = {}
        for mode, amount in ((FedRepTrainMode.HEAD, budget.head), (FedRepTrainMode.REPRESENTATION, budget.representation)):
            _, label, prefix = _PHASES[mode]
            self._enter_phase(mode)
            log(INFO, f"Beginning FedRep {label} Training Phase for {amount} {budget.unit.capitalize()}")
            phase_losses, phase_metrics = driver(amount, current_round)
            self._prefix_loss_and_metrics_dictionaries(prefix, phase_losses, phase_metrics)
            losses.update(phase_losses)
            metrics.update(phase_metrics)
        return losses, metrics

    def train_fedrep_by_epochs(self, head_epochs: int, rep_epochs: int, current_round: int | None = None) -> tuple[dict[str, float], dict[str, Scalar]]:
        return self._run_phases(_PhaseBudget("epochs", head_epochs, rep_epochs), current_round)

    def train_fedrep_by_steps(self, head_steps: int, rep_steps: int, current_round: int | None = None) -> tuple[dict[str, float], dict[str, Scalar]]:
        return self._run_phases(_PhaseBudget("steps", head_steps, rep_steps), current_round)

    # ------------------------------------------------------------------------------------------ protocol
    def fit(self, parameters: NDArrays, config: Config) -> tuple[NDArrays, int, dict[str, Scalar]]:
        round_began = datetime.datetime.now()
        budget = _PhaseBudget.from_config(config)
        _, server_round, evaluate_after_fit = self.process_fed_rep_config(config)
        if not self.initialized:
            self.setup_client(config)
        self.set_parameters(parameters, config, fitting_round=True)
        self.update_before_train(server_round)
        if not (budget.head and budget.representation):
            raise ValueError(f"Local epochs or steps not correctly specified: {budget.legacy_tuple()}")
        training_began = datetime.datetime.now()
        loss_dict, metrics = self._run_phases(budget, server_round)
        training_took = datetime.datetime.now() - training_began
        if self._should_evaluate_after_fit(evaluate_after_fit):
            validation_loss, validation_metrics = self.validate()
            metrics.update(validation_metrics)
            self._maybe_checkpoint(validation_loss, validation_metrics, CheckpointMode.PRE_AGGREGATION)
        self.reports_manager.report({"fit_metrics": metrics, "fit_losses": loss_dict, "round": server_round,
                                     "round_start": str(round_began), "fit_time_elapsed": str(training_took)}, server_round)
        return self.get_parameters(config), self.num_train_samples, metrics

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        if self.fedrep_train_mode not in _PHASES:
            raise ValueError("Training Mode in an invalid state")
        for optimizer in self.optimizers.values():
            optimizer.zero_grad()
        with self._amp():
            preds, features = self.predict(input)
            losses = self.compute_training_loss(preds, features, self.transform_target(target))
        losses.backward["backward"].backward()
        self.optimizers[_PHASES[self.fedrep_train_mode][0]].step()  # only the active half moves
        return losses, preds
