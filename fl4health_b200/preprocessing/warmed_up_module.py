"""Warm-starting a model from a (possibly differently structured) pretrained one (parity:
``fl4health/preprocessing/warmed_up_module.py:11-123``).

An optional JSON mapping ``{target_prefix: pretrained_prefix}`` redirects dotted key prefixes; without it keys are
matched by name.  Only entries whose shapes agree are loaded; everything else keeps its initial value."""

from __future__ import annotations

import json
import os
from logging import INFO, WARNING
from pathlib import Path

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.parallel.arena import arena_of


class WarmedUpModule:
    def __init__(
        self, pretrained_model: torch.nn.Module | None = None, pretrained_model_path: Path | None = None,
        weights_mapping_path: Path | None = None,
    ) -> None:
        if pretrained_model is not None and pretrained_model_path is not None:
            raise AssertionError("pretrained_model_path and pretrained_model is mutually exclusive. Please provide one of them.")
        if pretrained_model is not None:
            log(INFO, "Pretrained model is provided.")
            self.pretrained_model_state = pretrained_model.state_dict()
        elif pretrained_model_path is not None:
            assert os.path.exists(pretrained_model_path), f"Pretrained model path {pretrained_model_path} does not exist."
            log(INFO, f"Loading pretrained model from {pretrained_model_path}")
            self.pretrained_model_state = torch.load(pretrained_model_path, weights_only=False).state_dict()
        else:
            raise AssertionError("At least one of pretrained_model_path and pretrained_model should be provided.")
        self.weights_mapping_dict: dict[str, str] | None = None
        if weights_mapping_path is not None:
            with open(weights_mapping_path) as handle:
                self.weights_mapping_dict = json.load(handle)
        else:
            log(INFO, "Weights mapping dict is not provided. Matching states directly, based on target model's keys.")

    def get_matching_component(self, key: str) -> str | None:
        """Pretrained-model key for target key ``key`` (shortest mapped dotted prefix wins), or None if unmapped."""
        if self.weights_mapping_dict is None:
            return key
        parts = key.split(".")
        for depth in range(1, len(parts) + 1):
            prefix = ".".join(parts[:depth])
            if prefix in self.weights_mapping_dict:
                return self.weights_mapping_dict[prefix] + key[len(prefix):]
        return None

    def load_from_pretrained(self, model: torch.nn.Module) -> torch.nn.Module:
        target_state = model.state_dict()
        matched = {}
        for key, original in target_state.items():
            source_key = self.get_matching_component(key)
            if source_key is None:
                continue
            if source_key not in self.pretrained_model_state:
                log(WARNING, f"state won't be loaded. Key {source_key} not found in the pretrained model states.")
                continue
            candidate = self.pretrained_model_state[source_key]
            if candidate.size() != original.size():
                log(WARNING, f"State won't be loaded. Mismatched sizes {tuple(original.size())} -> ({source_key}) {tuple(candidate.size())}.")
                continue
            matched[key] = candidate
        log(INFO, f"{len(matched)}/{len(target_state)} states were matched.")
        arena = arena_of(model)
        if arena is not None:  # in-place into the flat arena (load_state_dict would also work; this avoids the key walk)
            with torch.no_grad():
                for key, value in matched.items():
                    target_state[key].copy_(value.to(target_state[key].dtype))
            arena.refresh_shadow()
            return model
        target_state.update(matched)
        model.load_state_dict(target_state)
        return model
