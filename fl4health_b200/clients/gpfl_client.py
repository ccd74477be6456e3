"""GPFL client (parity: ``fl4health/clients/gpfl_client.py:23-383``): three optimizers (``model``, ``gce``, ``cov``),
a frozen per-round GCE copy, conditional inputs built from the class-embedding table, and the loss
``CE + GCE-softmax + lam * ||global_features - embedding(target)||_2``.  GCE/CoV weight decay is forced to ``mu``."""

from __future__ import annotations

from logging import INFO, WARNING
from typing import Any

import torch
from torch.nn.functional import one_hot
from torch.optim import Optimizer

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config
from fl4health_b200.model_bases.gpfl_base import Gce, GpflModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.client import clone_and_freeze_model
from fl4health_b200.utils.losses import EvaluationLosses, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


_PARTS = ("model", "gce", "cov")  # one optimizer per trainable part
_REGULARISED = ("gce", "cov")      # weight decay of these is the algorithm's mu


class GpflClient(BasicClient):
    model: GpflModel
    gce_frozen: Gce
    feature_dim: int
    num_classes: int
    class_sample_proportion: torch.Tensor

    def __init__(self, *args: Any, lam: float = 0.01, mu: float = 0.01, **kwargs: Any) -> None:
        """``lam`` weights the magnitude-level loss, ``mu`` is the GCE / CoV weight decay; other arguments are
        ``BasicClient``'s."""
        super().__init__(*args, **kwargs)
        self.lam, self.mu = lam, mu
        if mu == 0.0:
            log(WARNING, "Mu parameter is set to 0.0: the GCE and CoV modules will not be regularized.")

    def _graph_variant(self) -> object:
        return id(getattr(self, "gce_frozen", None))  # the frozen GCE copy is rebuilt every round

    # ------------------------------------------------------------------------------------------ wiring
    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError("Return optimizers keyed 'model', 'gce' and 'cov'.")

    def set_optimizer(self, config: Config) -> None:
        per_part = self.get_optimizer(config)
        assert isinstance(per_part, dict) and set(per_part) == set(_PARTS), (
            f"Three optimizers must be defined with keys 'model', 'gce', and 'cov'; got {list(per_part)}")
        for part in _REGULARISED:
            groups = per_part[part].param_groups
            if any(group.get("weight_decay", 0.0) for group in groups):
                log(WARNING, f"Your {part} optimizer weight decay will be overwritten by the mu parameter.")
            for group in groups:
                group["weight_decay"] = self.mu
        log(INFO, f"GCE and CoV optimizer weight decay set to mu = {self.mu}")
        self.optimizers = per_part

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, GpflModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self.num_classes, self.feature_dim = self.model.num_classes, self.model.feature_dim
        self.class_sample_proportion = self.calculate_class_sample_proportions()

    # ------------------------------------------------------------------------------------------ conditioning
    def calculate_class_sample_proportions(self) -> torch.Tensor:
        """Empirical label distribution of the local training set (hard or one-hot / soft labels)."""
        histogram = torch.zeros(self.num_classes, device=self.device)
        for _, labels in self.train_loader:
            labels = labels.to(self.device)
            if labels.dim() != 2:
                labels = one_hot(labels.long(), num_classes=self.num_classes)
            assert labels.shape[1] == self.num_classes
            histogram += labels.sum(0)
        return histogram / histogram.sum()

    def compute_conditional_inputs(self) -> None:
        """Global context = mean class prototype; personalised context = prototypes weighted by the local label mix."""
        prototypes = self.gce_frozen.embedding.weight
        self.global_conditional_input = prototypes.mean(dim=0)
        self.personalized_conditional_input = (prototypes.T @ self.class_sample_proportion) / self.num_classes

    def update_before_train(self, current_server_round: int) -> None:
        frozen = clone_and_freeze_model(self.model.gce)
        assert isinstance(frozen, Gce)
        self.gce_frozen = frozen
        self.compute_conditional_inputs()
        super().update_before_train(current_server_round)

    def transform_input(self, input: TorchInputType) -> TorchInputType:
        conditioned = {"input": input} if isinstance(input, torch.Tensor) else input
        conditioned["global_conditional_input"] = self.global_conditional_input.detach()
        conditioned["personalized_conditional_input"] = self.personalized_conditional_input.detach()
        return conditioned

    # ------------------------------------------------------------------------------------------ step
    def compute_magnitude_level_loss(self, global_features: torch.Tensor, target: TorchTargetType) -> torch.Tensor:
        assert isinstance(target, torch.Tensor), "GPFL clients take only tensor targets."
        return torch.linalg.vector_norm(global_features - self.gce_frozen.lookup(target).detach())

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        task, _ = self.compute_loss_and_additional_losses(preds, features, target)
        angle = self.model.gce(features["global_features"], target)
        magnitude = self.compute_magnitude_level_loss(features["global_features"], target)
        recorded = {"prediction_loss": task.clone(), "gce_softmax_loss": angle.clone(), "magnitude_level_loss": magnitude.clone()}
        return TrainingLosses(backward=task + angle + self.lam * magnitude, additional_losses=recorded)

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        steppers = [self.optimizers[part] for part in _PARTS]
        for optimizer in steppers:
            optimizer.zero_grad()
        with self._amp():
            preds, features = self.predict(self.transform_input(input))
            losses = self.compute_training_loss(preds, features, self.transform_target(target))
        losses.backward["backward"].backward()
        self.transform_gradients(losses)
        for optimizer in steppers:
            optimizer.step()
        return losses, preds

    def val_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
        return super().val_step(self.transform_input(input), target)
