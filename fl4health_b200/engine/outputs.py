"""One place that runs a network on a client batch and normalises what came back.

Client models may take a tensor or keyword tensors and may return a tensor, a dict of predictions, or a
``(predictions, features)`` pair; every client flavour (``BasicClient.predict``, the model-parameterised
``FlexibleClient.predict_with_model``, Ditto's twin forward) needs the same two steps."""

from __future__ import annotations

from typing import Any

import torch
from torch import nn

PAIR = 2  # a tuple result must be (predictions, features)


def call_model(model: nn.Module, batch: Any) -> Any:
    if isinstance(batch, dict):
        return model(**batch)
    if isinstance(batch, torch.Tensor):
        return model(batch)
    raise TypeError('"input" must be of type torch.Tensor or dict[str, torch.Tensor].')


def as_preds_and_features(output: Any) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
    if isinstance(output, torch.Tensor):
        return {"prediction": output}, {}
    if isinstance(output, dict):
        return output, {}
    if not isinstance(output, tuple):
        raise ValueError("Model forward did not return a tensor, dictionary of tensors, or tuple of tensors")
    if len(output) != PAIR:
        raise ValueError(f"Output tuple should have length 2 but has length {len(output)}")
    return output[0], output[1]


def forward(model: nn.Module, batch: Any) -> tuple[dict[str, torch.Tensor], dict[str, torch.Tensor]]:
    return as_preds_and_features(call_model(model, batch))
