"""GPFL building blocks (Zhang et al. 2023; parity: ``fl4health/model_bases/gpfl_base.py:12-278``):
``Gce`` — global category embedding with a cosine-softmax loss; ``CoV`` — conditional affine modulation
``relu(f * (gamma(ctx) + 1) + beta(ctx))``; ``GpflModel`` = base -> CoV -> head (+ global-context features in
training).  Exchanged: base module, CoV and GCE."""

from __future__ import annotations

from logging import WARNING

import torch
import torch.nn.functional as F
from torch import nn

from fl4health_b200.common.logger import log
from fl4health_b200.model_bases.partial_layer_exchange_model import PartialLayerExchangeModel
from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel


def _class_indices_or_distribution(label: torch.Tensor, num_classes: int, dtype: torch.dtype) -> torch.Tensor:
    """Hard labels stay indices, ``[B, num_classes]`` rows are taken as target distributions (one-hot included)."""
    if label.dim() == 1:
        return label.long()
    assert label.shape[1] == num_classes, "One-hot labels must have shape (batch_size, num_classes)."
    return label.to(dtype)


class Gce(nn.Module):
    """Global category embedding: one learnable prototype per class, trained with a softmax over cosine similarities
    between the (normalised) features and the (normalised) prototypes."""

    def __init__(self, feature_dim: int, num_classes: int) -> None:
        super().__init__()
        self.feature_dim, self.num_classes = feature_dim, num_classes
        self.embedding = nn.Embedding(num_classes, feature_dim)

    def forward(self, feature_tensor: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        similarities = F.normalize(feature_tensor) @ F.normalize(self.embedding.weight).T
        return F.cross_entropy(similarities, _class_indices_or_distribution(label, self.num_classes, similarities.dtype))

    def lookup(self, target: torch.Tensor) -> torch.Tensor:
        if self.training:
            log(WARNING, "Lookup is an embedding read-out (no forward pass) and is not meant for training mode.")
        indices = _class_indices_or_distribution(target, self.num_classes, torch.float32)
        if indices.dim() == 2:
            indices = indices.argmax(dim=1)
        assert indices.dim() == 1, "lookup requires 1D tensor of class indices."
        return self.embedding.weight.data[indices.long()]


def _conditioner(width: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(width, width), nn.ReLU(), nn.LayerNorm([width]))


class CoV(nn.Module):
    """Conditional valve: ``relu(f * (gamma(context) + 1) + beta(context))``."""

    def __init__(self, feature_dim: int) -> None:
        super().__init__()
        self.conditional_gamma, self.conditional_beta = _conditioner(feature_dim), _conditioner(feature_dim)
        self.activation = nn.ReLU()

    def forward(self, feature_tensor: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        scale = self.conditional_gamma(context) + 1
        return self.activation(torch.addcmul(self.conditional_beta(context), feature_tensor, scale))


class GpflBaseAndHeadModules(SequentiallySplitExchangeBaseModel):
    def forward(self, input: torch.Tensor) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        raise NotImplementedError("Use GpflModel.forward; this container only groups base and head.")


class GpflModel(PartialLayerExchangeModel):
    _SHARED_PREFIXES = ("cov.", "gce.")  # exchanged next to the base feature extractor

    def __init__(self, base_module: nn.Module, head_module: nn.Module, feature_dim: int, num_classes: int,
                 flatten_features: bool = False) -> None:
        super().__init__()
        self.feature_dim, self.num_classes = feature_dim, num_classes
        self.gpfl_main_module = GpflBaseAndHeadModules(base_module, head_module, flatten_features)
        self.cov = CoV(feature_dim)
        self.gce = Gce(feature_dim, num_classes)

    def forward(
        self, input: torch.Tensor, global_conditional_input: torch.Tensor, personalized_conditional_input: torch.Tensor
    ) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        features = self.gpfl_main_module.features_forward(input)
        assert features.shape[1] == self.feature_dim, "Base-module output width must equal feature_dim."
        personalised = self.cov(features, personalized_conditional_input)
        preds = {"prediction": self.gpfl_main_module.head_module(personalised)}
        if not self.training:
            return preds, {}
        assert len(global_conditional_input) == self.feature_dim
        return preds, {"local_features": personalised, "global_features": self.cov(features, global_conditional_input)}

    def layers_to_exchange(self) -> list[str]:
        shared = [key for key in self.state_dict() if key.startswith(self._SHARED_PREFIXES)]
        return [f"gpfl_main_module.{key}" for key in self.gpfl_main_module.layers_to_exchange()] + shared
