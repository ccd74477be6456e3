"""GPFL client (parity: ``fl4health/clients/gpfl_client.py:23-383``): three optimizers (``model``, ``gce``, ``cov``),
a frozen per-round GCE copy, conditional inputs built from the class-embedding table, and the loss
``CE + GCE-softmax + lam * ||global_features - embedding(target)||_2``.  GCE/CoV weight decay is forced to ``mu``."""

from __future__ import annotations

from collections.abc import Sequence
from logging import INFO, WARNING
from pathlib import Path
from typing import Any

import torch
from torch.nn.functional import one_hot
from torch.optim import Optimizer

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config
from fl4health_b200.metrics.base_metrics import Metric
from fl4health_b200.model_bases.gpfl_base import Gce, GpflModel
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.utils.client import clone_and_freeze_model
from fl4health_b200.utils.losses import EvaluationLosses, LossMeterType, TrainingLosses
from fl4health_b200.utils.typing import TorchFeatureType, TorchInputType, TorchPredType, TorchTargetType


class GpflClient(BasicClient):
    def __init__(
        self,
        data_path: Path,
        metrics: Sequence[Metric],
        device: torch.device,
        loss_meter_type: LossMeterType = LossMeterType.AVERAGE,
        checkpoint_and_state_module: Any = None,
        reporters: Any = None,
        progress_bar: bool = False,
        client_name: str | None = None,
        lam: float = 0.01,
        mu: float = 0.01,
        engine_options: Any = None,
    ) -> None:
        super().__init__(data_path=data_path, metrics=metrics, device=device, loss_meter_type=loss_meter_type,
                         checkpoint_and_state_module=checkpoint_and_state_module, reporters=reporters,
                         progress_bar=progress_bar, client_name=client_name, engine_options=engine_options)
        self.lam, self.mu = lam, mu
        self.model: GpflModel
        self.gce_frozen: Gce
        self.feature_dim: int
        self.num_classes: int
        self.class_sample_proportion: torch.Tensor
        if mu == 0.0:
            log(WARNING, "Mu parameter is set to 0.0: the GCE and CoV modules will not be regularized.")

    def _graph_variant(self) -> object:
        return id(getattr(self, "gce_frozen", None))

    def get_optimizer(self, config: Config) -> dict[str, Optimizer]:
        raise NotImplementedError("Return optimizers keyed 'model', 'gce' and 'cov'.")

    def set_optimizer(self, config: Config) -> None:
        optimizers = self.get_optimizer(config)
        assert isinstance(optimizers, dict) and set(optimizers.keys()) == {"model", "gce", "cov"}, (
            f"Three optimizers must be defined with keys 'model', 'gce', and 'cov'; got {list(optimizers)}"
        )
        for key in ("gce", "cov"):
            if any(group.get("weight_decay", 0.0) != 0.0 for group in optimizers[key].param_groups):
                log(WARNING, f"Your {key} optimizer weight decay will be overwritten by the mu parameter.")
            for group in optimizers[key].param_groups:
                group["weight_decay"] = self.mu
        log(INFO, f"GCE and CoV optimizer weight decay set to mu = {self.mu}")
        self.optimizers = optimizers

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        assert isinstance(self.model, GpflModel)
        return FixedLayerExchanger(self.model.layers_to_exchange())

    def calculate_class_sample_proportions(self) -> torch.Tensor:
        counts = torch.zeros(self.num_classes, device=self.device)
        for _, target in self.train_loader:
            target = target.to(self.device)
            if target.dim() == 2:
                assert target.shape[1] == self.num_classes
                counts += target.sum(0)
            else:
                counts += one_hot(target.long(), num_classes=self.num_classes).sum(0)
        return counts / counts.sum()

    def setup_client(self, config: Config) -> None:
        super().setup_client(config)
        self.num_classes, self.feature_dim = self.model.num_classes, self.model.feature_dim
        self.class_sample_proportion = self.calculate_class_sample_proportions()

    def compute_conditional_inputs(self) -> None:
        embeddings = self.gce_frozen.embedding.weight
        self.global_conditional_input = embeddings.sum(0) / self.num_classes
        self.personalized_conditional_input = torch.matmul(embeddings.T, self.class_sample_proportion) / self.num_classes

    def update_before_train(self, current_server_round: int) -> None:
        frozen = clone_and_freeze_model(self.model.gce)
        assert isinstance(frozen, Gce)
        self.gce_frozen = frozen
        self.compute_conditional_inputs()
        return super().update_before_train(current_server_round)

    def transform_input(self, input: TorchInputType) -> TorchInputType:
        extras = {"global_conditional_input": self.global_conditional_input.detach(),
                  "personalized_conditional_input": self.personalized_conditional_input.detach()}
        if isinstance(input, torch.Tensor):
            return {"input": input, **extras}
        input.update(extras)
        return input

    def train_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[TrainingLosses, TorchPredType]:
        for key in ("model", "gce", "cov"):
            self.optimizers[key].zero_grad()
        with self._amp():
            preds, features = self.predict(self.transform_input(input))
            target = self.transform_target(target)
            losses = self.compute_training_loss(preds, features, target)
        losses.backward["backward"].backward()
        self.transform_gradients(losses)
        for key in ("model", "gce", "cov"):
            self.optimizers[key].step()
        return losses, preds

    def compute_magnitude_level_loss(self, global_features: torch.Tensor, target: TorchTargetType) -> torch.Tensor:
        assert isinstance(target, torch.Tensor), "GPFL clients take only tensor targets."
        return torch.norm(global_features - self.gce_frozen.lookup(target).detach(), 2)

    def compute_training_loss(self, preds: TorchPredType, features: TorchFeatureType, target: TorchTargetType) -> TrainingLosses:
        prediction_loss, _ = self.compute_loss_and_additional_losses(preds, features, target)
        gce_softmax_loss = self.model.gce(features["global_features"], target)
        magnitude_level_loss = self.compute_magnitude_level_loss(features["global_features"], target)
        loss = prediction_loss + gce_softmax_loss + magnitude_level_loss * self.lam
        return TrainingLosses(backward=loss, additional_losses={
            "prediction_loss": prediction_loss.clone(), "gce_softmax_loss": gce_softmax_loss.clone(),
            "magnitude_level_loss": magnitude_level_loss.clone()})

    def val_step(self, input: TorchInputType, target: TorchTargetType) -> tuple[EvaluationLosses, TorchPredType]:
        return super().val_step(self.transform_input(input), target)
