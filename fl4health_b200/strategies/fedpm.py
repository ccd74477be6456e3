"""FedPM server strategy (Isik et al. 2023; parity: ``fl4health/strategies/fedpm.py:12-162``).

Clients upload Bernoulli-sampled binary masks.  Aggregation is the uniform mean, or — ``bayesian_aggregation`` — a
Beta posterior per score: ``alpha += sum M``, ``beta += K - sum M``, ``theta = (alpha - 1) / (alpha + beta - 2)``.
The vote runs in one kernel per tensor over uint8 masks (``ops.flat.fedpm_vote``)."""

from __future__ import annotations

from collections import defaultdict
from typing import Any

import torch

from fl4health_b200.common.typing import NDArray, NDArrays, to_tensor
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.strategies.fedavg_dynamic_layer import FedAvgDynamicLayer


class FedPm(FedAvgDynamicLayer):
    def __init__(self, *, bayesian_aggregation: bool = True, **kwargs: Any) -> None:
        kwargs["weighted_aggregation"] = False
        super().__init__(**kwargs)
        self.beta_parameters: dict[str, tuple[torch.Tensor, torch.Tensor]] = {}
        self.bayesian_aggregation = bayesian_aggregation

    def aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, NDArray]:
        if not self.bayesian_aggregation:
            return super().aggregate(results)
        return self.aggregate_bayesian(results)

    def aggregate_bayesian(self, results: list[tuple[NDArrays, int]]) -> dict[str, NDArray]:
        names_to_masks: defaultdict[str, list[torch.Tensor]] = defaultdict(list)
        for packed_layers, _ in results:
            layers, names = self.parameter_packer.unpack_parameters(packed_layers)
            for layer, name in zip(layers, names):
                mask = to_tensor(layer)
                names_to_masks[name].append(mask)
                if name not in self.beta_parameters:
                    ones = torch.ones(mask.shape, dtype=torch.float32, device=mask.device)
                    self.beta_parameters[name] = (ones, ones.clone())
        out: dict[str, NDArray] = {}
        for name, (alpha, beta) in self.beta_parameters.items():
            masks = names_to_masks.get(name)
            if not masks:
                continue
            flat_masks = [m.to(device=alpha.device, dtype=torch.uint8).reshape(-1).contiguous() for m in masks]
            theta = flat_ops.fedpm_vote(flat_masks, alpha.reshape(-1), beta.reshape(-1), bayesian=True)
            out[name] = theta.reshape(alpha.shape)
        return out

    def reset_beta_priors(self) -> None:
        """Forget accumulated evidence (called every ``reset_frequency`` rounds by ``FedPmServer``)."""
        for name, (alpha, beta) in self.beta_parameters.items():
            self.beta_parameters[name] = (torch.ones_like(alpha), torch.ones_like(beta))
