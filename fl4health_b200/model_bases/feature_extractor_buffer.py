"""Forward-hook feature capture by layer-name prefix (parity: ``fl4health/model_bases/feature_extractor_buffer.py:10-182``).

Used by the MK-MMD / Deep-MMD clients to read intermediate activations of both the personal and the global model.
Features stay on the device; nothing is copied until ``get_extracted_features`` concatenates them.
"""

from __future__ import annotations

from collections.abc import Callable
from logging import INFO

import torch
from torch import nn
from torch.utils.hooks import RemovableHandle

from fl4health_b200.common.logger import log


class _LayerTap:
    """One tapped layer: the forward hook, what it has captured since the last clear, and how it is handed out."""

    def __init__(self, flatten: bool) -> None:
        self.flatten = flatten
        self.captured: list[torch.Tensor] = []
        self.handle: RemovableHandle | None = None

    def attach(self, module: nn.Module, keep_history: Callable[[], bool]) -> None:
        def on_forward(_module: nn.Module, _inputs: tuple, output: torch.Tensor) -> None:
            if keep_history():
                self.captured.append(output)
            else:
                self.captured = [output]

        self.handle = module.register_forward_hook(on_forward)

    def detach(self) -> None:
        if self.handle is not None:
            self.handle.remove()
            self.handle = None

    def features(self) -> torch.Tensor:
        joined = self.captured[0] if len(self.captured) == 1 else torch.cat(self.captured, dim=0)
        return joined.reshape(len(joined), -1) if self.flatten else joined


class FeatureExtractorBuffer:
    def __init__(self, model: nn.Module, flatten_feature_extraction_layers: dict[str, bool]) -> None:
        """``flatten_feature_extraction_layers`` maps a layer-name *prefix* to whether its output is flattened to 2-D.
        The hook lands on the LAST named module whose name starts with the prefix (module order = forward order)."""
        self.model = model
        self.flatten_feature_extraction_layers = flatten_feature_extraction_layers
        self.accumulate_features = False
        self._taps = {layer: _LayerTap(flatten) for layer, flatten in flatten_feature_extraction_layers.items()}

    # the reference's attribute views of the same state
    @property
    def fhooks(self) -> list[RemovableHandle]:
        return [tap.handle for tap in self._taps.values() if tap.handle is not None]

    @property
    def extracted_features_buffers(self) -> dict[str, list[torch.Tensor]]:
        return {layer: tap.captured for layer, tap in self._taps.items()}

    def enable_accumulating_features(self) -> None:
        self.accumulate_features = True

    def disable_accumulating_features(self) -> None:
        self.accumulate_features = False

    def clear_buffers(self) -> None:
        for tap in self._taps.values():
            tap.captured = []

    def get_hierarchical_attr(self, module: nn.Module, layer_hierarchy: list[str]) -> nn.Module:
        for part in layer_hierarchy:
            module = getattr(module, part)
        return module

    def find_last_common_prefix(self, prefix: str, layers_name: list[str]) -> str:
        candidates = [name for name in layers_name if name.startswith(prefix)]
        if not candidates:
            raise ValueError(f"no module of the model starts with '{prefix}'")
        return candidates[-1]

    def _maybe_register_hooks(self) -> None:
        if self.fhooks:
            log(INFO, "Hooks already registered.")
            return
        log(INFO, "Starting to register hooks:")
        module_names = [name for name, _ in self.model.named_modules()]
        for layer, tap in self._taps.items():
            log(INFO, f"Registering hook for layer: {layer}")
            path = self.find_last_common_prefix(layer, module_names).split(".")
            tap.attach(self.get_hierarchical_attr(self.model, path), keep_history=lambda: self.accumulate_features)

    def remove_hooks(self) -> None:
        """Hooks hold closures and make the module unpicklable: remove before checkpointing."""
        log(INFO, "Removing hooks.")
        for tap in self._taps.values():
            tap.detach()

    def forward_hook(self, layer_name: str) -> Callable:
        """A hook function for ``layer_name`` (kept for callers that register hooks themselves)."""
        tap = self._taps[layer_name]

        def hook(module: nn.Module, input: torch.Tensor, output: torch.Tensor) -> None:  # noqa: ARG001
            tap.captured = [*tap.captured, output] if self.accumulate_features else [output]

        return hook

    def flatten(self, features: torch.Tensor) -> torch.Tensor:
        return features.reshape(len(features), -1)

    def get_extracted_features(self) -> dict[str, torch.Tensor]:
        return {layer: tap.features() for layer, tap in self._taps.items()}
