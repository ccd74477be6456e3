"""FedRep model: freeze/unfreeze the representation or the head (parity: ``fedrep_base.py:4-32``)."""

from __future__ import annotations

from torch import nn

from fl4health_b200.model_bases.sequential_split_models import SequentiallySplitExchangeBaseModel


class FedRepModel(SequentiallySplitExchangeBaseModel):
    def __init__(self, base_module: nn.Module, head_module: nn.Module, flatten_features: bool = False) -> None:
        super().__init__(base_module, head_module, flatten_features)

    @staticmethod
    def _set_trainable(module: nn.Module, flag: bool) -> None:
        for param in module.parameters():
            param.requires_grad = flag

    def freeze_base_module(self) -> None:
        self._set_trainable(self.base_module, False)

    def unfreeze_base_module(self) -> None:
        self._set_trainable(self.base_module, True)

    def freeze_head_module(self) -> None:
        self._set_trainable(self.head_module, False)

    def unfreeze_head_module(self) -> None:
        self._set_trainable(self.head_module, True)
