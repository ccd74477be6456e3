"""APFL twin model: ``personal = alpha * local + (1 - alpha) * global`` (parity: ``apfl_base.py:9-129``).

``update_alpha`` needs ``sum_l <w_loc - w_glob, alpha g_loc + (1-alpha) g_glob>``.  The reference accumulates it with a
``.cpu().numpy().item()`` host sync per layer (:98-116); here it is one flat reduction kernel when the module lives in
an arena (local and global sub-models have identical layouts, so they are two equal-length ranges of the arena), and a
single on-device stack otherwise — one sync either way.
"""

from __future__ import annotations

import copy

import torch
from torch import nn

from fl4health_b200.model_bases.partial_layer_exchange_model import PartialLayerExchangeModel


class ApflModule(PartialLayerExchangeModel):
    def __init__(self, model: nn.Module, adaptive_alpha: bool = True, alpha: float = 0.5, alpha_lr: float = 0.01) -> None:
        super().__init__()
        self.local_model: nn.Module = model
        self.global_model: nn.Module = copy.deepcopy(model)
        self.adaptive_alpha = adaptive_alpha
        self.alpha = alpha
        self.alpha_lr = alpha_lr

    def global_forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.global_model(input)

    def local_forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.local_model(input)

    def forward(self, input: torch.Tensor) -> dict[str, torch.Tensor]:
        global_logits = self.global_forward(input)
        local_logits = self.local_forward(input)
        personal_logits = self.alpha * local_logits + (1.0 - self.alpha) * global_logits
        return {"personal": personal_logits, "global": global_logits, "local": local_logits}

    def _alpha_gradient(self) -> float:
        from fl4health_b200.ops import flat as flat_ops
        from fl4health_b200.parallel.arena import arena_of

        local_params = [p for p in self.local_model.parameters() if p.requires_grad]
        global_params = [p for p in self.global_model.parameters() if p.requires_grad]
        arena = arena_of(self)
        if arena is not None and arena.grad is not None and local_params:
            loc_names = [n for n, p in self.local_model.named_parameters() if p.requires_grad]
            loc = arena.range_of([f"local_model.{n}" for n in loc_names])
            glob = arena.range_of([f"global_model.{n}" for n in loc_names])
            if len(loc) == 1 and len(glob) == 1 and loc[0][1] - loc[0][0] == glob[0][1] - glob[0][0]:
                (ls, le), (gs, ge) = loc[0], glob[0]
                return float(flat_ops.apfl_alpha_grad(arena.flat[ls:le], arena.flat[gs:ge], arena.grad[ls:le],
                                                      arena.grad[gs:ge], float(self.alpha)).item())
        terms = []
        for local_p, global_p in zip(local_params, global_params):
            assert local_p.grad is not None and global_p.grad is not None
            mixed = self.alpha * local_p.grad + (1.0 - self.alpha) * global_p.grad
            terms.append(((local_p - global_p) * mixed).sum())
        return float(torch.stack(terms).sum().item()) if terms else 0.0

    def update_alpha(self) -> None:
        """alpha <- clip(alpha - alpha_lr * (d/d alpha + 0.02 alpha), 0, 1)  (as in the APFL reference implementation)."""
        grad_alpha = self._alpha_gradient() + 0.02 * self.alpha
        self.alpha = max(min(self.alpha - self.alpha_lr * grad_alpha, 1.0), 0.0)

    def layers_to_exchange(self) -> list[str]:
        return [name for name in self.state_dict() if name.startswith("global_model.")]
