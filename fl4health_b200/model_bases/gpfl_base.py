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


class Gce(nn.Module):
    def __init__(self, feature_dim: int, num_classes: int) -> None:
        super().__init__()
        self.feature_dim = feature_dim
        self.num_classes = num_classes
        self.embedding = nn.Embedding(num_classes, feature_dim)

    def forward(self, feature_tensor: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        cosine = F.linear(F.normalize(feature_tensor), F.normalize(self.embedding.weight))
        if label.dim() == 1:
            one_hot = F.one_hot(label.long(), self.num_classes).to(cosine.dtype)
        else:
            assert label.shape[1] == self.num_classes, "One-hot labels must have shape (batch_size, num_classes)."
            one_hot = label.to(cosine.dtype)
        return -(one_hot * F.log_softmax(cosine, dim=1)).sum(dim=1).mean()

    def lookup(self, target: torch.Tensor) -> torch.Tensor:
        if self.training:
            log(WARNING, "Lookup is an embedding read-out (no forward pass) and is not meant for training mode.")
        if target.dim() == 2:
            assert target.shape[1] == self.num_classes, "One-hot labels must have shape (batch_size, num_classes)."
            target = torch.argmax(target, dim=1)
        assert target.dim() == 1, "lookup requires 1D tensor of class indices."
        return self.embedding.weight.data[target.long()]


class CoV(nn.Module):
    def __init__(self, feature_dim: int) -> None:
        super().__init__()
        self.conditional_gamma = nn.Sequential(nn.Linear(feature_dim, feature_dim), nn.ReLU(), nn.LayerNorm([feature_dim]))
        self.conditional_beta = nn.Sequential(nn.Linear(feature_dim, feature_dim), nn.ReLU(), nn.LayerNorm([feature_dim]))
        self.activation = nn.ReLU()

    def forward(self, feature_tensor: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        gamma, beta = self.conditional_gamma(context), self.conditional_beta(context)
        return self.activation(feature_tensor * (gamma + 1) + beta)


class GpflBaseAndHeadModules(SequentiallySplitExchangeBaseModel):
    def forward(self, input: torch.Tensor) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        raise NotImplementedError("Use GpflModel.forward; this container only groups base and head.")


class GpflModel(PartialLayerExchangeModel):
    def __init__(self, base_module: nn.Module, head_module: nn.Module, feature_dim: int, num_classes: int,
                 flatten_features: bool = False) -> None:
        super().__init__()
        self.feature_dim = feature_dim
        self.num_classes = num_classes
        self.gpfl_main_module = GpflBaseAndHeadModules(base_module, head_module, flatten_features)
        self.cov = CoV(feature_dim)
        self.gce = Gce(feature_dim, num_classes)

    def forward(
        self, input: torch.Tensor, global_conditional_input: torch.Tensor, personalized_conditional_input: torch.Tensor
    ) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
        features = self.gpfl_main_module.features_forward(input)
        assert features.shape[1] == self.feature_dim, "Base-module output width must equal feature_dim."
        local_features = self.cov(features, personalized_conditional_input)
        predictions = self.gpfl_main_module.head_module(local_features)
        if not self.training:
            return {"prediction": predictions}, {}
        assert len(global_conditional_input) == self.feature_dim
        global_features = self.cov(features, global_conditional_input)
        return {"prediction": predictions}, {"local_features": local_features, "global_features": global_features}

    def layers_to_exchange(self) -> list[str]:
        base = [f"gpfl_main_module.{name}" for name in self.gpfl_main_module.layers_to_exchange()]
        return base + [name for name in self.state_dict() if name.startswith(("cov.", "gce."))]
