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


class FeatureExtractorBuffer:
    def __init__(self, model: nn.Module, flatten_feature_extraction_layers: dict[str, bool]) -> None:
        """``flatten_feature_extraction_layers`` maps a layer-name *prefix* to whether its output is flattened to 2-D.
        The hook lands on the LAST named module whose name starts with the prefix (module order = forward order)."""
        self.model = model
        self.flatten_feature_extraction_layers = flatten_feature_extraction_layers
        self.fhooks: list[RemovableHandle] = []
        self.accumulate_features = False
        self.extracted_features_buffers: dict[str, list[torch.Tensor]] = {}
        self.clear_buffers()

    def enable_accumulating_features(self) -> None:
        self.accumulate_features = True

    def disable_accumulating_features(self) -> None:
        self.accumulate_features = False

    def clear_buffers(self) -> None:
        self.extracted_features_buffers = {layer: [] for layer in self.flatten_feature_extraction_layers}

    def get_hierarchical_attr(self, module: nn.Module, layer_hierarchy: list[str]) -> nn.Module:
        for part in layer_hierarchy:
            module = getattr(module, part)
        return module

    def find_last_common_prefix(self, prefix: str, layers_name: list[str]) -> str:
        matches = [name for name in layers_name if name.startswith(prefix)]
        if not matches:
            raise ValueError(f"no module of the model starts with '{prefix}'")
        return matches[-1]

    def _maybe_register_hooks(self) -> None:
        if self.fhooks:
            log(INFO, "Hooks already registered.")
            return
        log(INFO, "Starting to register hooks:")
        names = [name for name, _ in self.model.named_modules()]
        for layer in self.flatten_feature_extraction_layers:
            log(INFO, f"Registering hook for layer: {layer}")
            target = self.get_hierarchical_attr(self.model, self.find_last_common_prefix(layer, names).split("."))
            self.fhooks.append(target.register_forward_hook(self.forward_hook(layer)))

    def remove_hooks(self) -> None:
        """Hooks hold closures and make the module unpicklable: remove before checkpointing."""
        log(INFO, "Removing hooks.")
        for hook in self.fhooks:
            hook.remove()
        self.fhooks.clear()

    def forward_hook(self, layer_name: str) -> Callable:
        def hook(module: nn.Module, input: torch.Tensor, output: torch.Tensor) -> None:  # noqa: ARG001
            if self.accumulate_features:
                self.extracted_features_buffers[layer_name].append(output)
            else:
                self.extracted_features_buffers[layer_name] = [output]

        return hook

    def flatten(self, features: torch.Tensor) -> torch.Tensor:
        return features.reshape(len(features), -1)

    def get_extracted_features(self) -> dict[str, torch.Tensor]:
        out = {}
        for layer, chunks in self.extracted_features_buffers.items():
            joined = chunks[0] if len(chunks) == 1 else torch.cat(chunks, dim=0)
            out[layer] = self.flatten(joined) if self.flatten_feature_extraction_layers[layer] else joined
        return out
