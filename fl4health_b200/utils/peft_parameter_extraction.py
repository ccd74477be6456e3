"""LoRA / PEFT adapter extraction (parity: ``fl4health/utils/peft_parameter_extraction.py:7-19``).

Uses ``peft.get_peft_model_state_dict`` when the optional ``peft`` package is installed; otherwise falls back to the
naming convention PEFT itself uses (adapter tensors carry ``lora_`` / ``modules_to_save`` / prompt-encoder markers in
their state-dict keys), so federating adapters does not hard-require the dependency."""

from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

from fl4health_b200.common.typing import NDArrays, Parameters, ndarrays_to_parameters

_ADAPTER_MARKERS = ("lora_", "modules_to_save", "prompt_encoder", "ia3_", "adapter_")


def get_peft_state_dict(model: nn.Module) -> OrderedDict[str, torch.Tensor]:
    try:
        from peft import get_peft_model_state_dict  # type: ignore[import-not-found]

        return OrderedDict(get_peft_model_state_dict(model))
    except ImportError:
        return OrderedDict((k, v) for k, v in model.state_dict().items() if any(m in k for m in _ADAPTER_MARKERS))


def get_all_peft_parameters_from_model(model: nn.Module) -> Parameters:
    return ndarrays_to_parameters(NDArrays([v.detach() for v in get_peft_state_dict(model).values()]))
