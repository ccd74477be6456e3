"""Layer-subset exchangers (parity: ``fl4health/parameter_exchange/layer_exchanger.py:17-155``).

With an arena-backed model a fixed layer subset is a set of offset ranges; the views handed out here are slices of
the rank's flat buffer and pulls are in-place copies into it.  The two fixed-subset exchangers differ only in how the
list of state-dict keys is produced, so they share ``_NamedSubsetExchanger``.
"""

from __future__ import annotations

from collections.abc import Set
from typing import TypeVar

from torch import nn

from fl4health_b200.common.typing import Config, NDArrays
from fl4health_b200.parameter_exchange._state import inject_state, state_views
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithLayerNames
from fl4health_b200.parameter_exchange.partial_parameter_exchanger import PartialParameterExchanger
from fl4health_b200.utils.typing import LayerSelectionFunction

TorchModule = TypeVar("TorchModule", bound=nn.Module)  # element type of ``module_exclusions``


class _NamedSubsetExchanger(ParameterExchanger):
    """Exchange exactly the state-dict entries named in ``layers_to_transfer`` (order = wire order)."""

    layers_to_transfer: list[str]

    def apply_layer_filter(self, model: nn.Module) -> NDArrays:
        return state_views(model, self.layers_to_transfer)

    def push_parameters(
        self, model: nn.Module, initial_model: nn.Module | None = None, config: Config | None = None
    ) -> NDArrays:
        return self.apply_layer_filter(model)

    def pull_parameters(self, parameters: NDArrays, model: nn.Module, config: Config | None = None) -> None:
        expected = len(self.layers_to_transfer)
        assert len(parameters) == expected, f"received {len(parameters)} arrays for {expected} exchanged layers"
        inject_state(model, self.layers_to_transfer, parameters)


class FixedLayerExchanger(_NamedSubsetExchanger):
    def __init__(self, layers_to_transfer: list[str]) -> None:
        self.layers_to_transfer = layers_to_transfer


class LayerExchangerWithExclusions(_NamedSubsetExchanger):
    """Exchange everything except state belonging to modules of the excluded *types* (FedBN)."""

    def __init__(self, model: nn.Module, module_exclusions: Set[type[nn.Module]]) -> None:
        self.module_exclusions = module_exclusions
        named = model.named_modules(remove_duplicate=False)
        self.modules_to_filter: set[str] = {path for path, module in named if path and self.should_module_be_excluded(module)}
        self.layers_to_transfer = self.get_layers_to_transfer(model)

    def should_module_be_excluded(self, module: nn.Module) -> bool:
        return type(module) in self.module_exclusions

    def should_layer_be_excluded(self, layer_name: str) -> bool:
        return layer_name.startswith(tuple(self.modules_to_filter)) if self.modules_to_filter else False

    def get_layers_to_transfer(self, model: nn.Module) -> list[str]:
        return [key for key in model.state_dict() if not self.should_layer_be_excluded(key)]


class DynamicLayerExchanger(PartialParameterExchanger[list[str]]):
    """Subset chosen per round by a ``LayerSelectionFunction``; names ride in the trailing slot."""

    def __init__(self, layer_selection_function: LayerSelectionFunction) -> None:
        self.layer_selection_function = layer_selection_function
        self.parameter_packer = ParameterPackerWithLayerNames()

    def select_parameters(self, model: nn.Module, initial_model: nn.Module | None = None) -> tuple[NDArrays, list[str]]:
        return self.layer_selection_function(model, initial_model)

    def push_parameters(
        self, model: nn.Module, initial_model: nn.Module | None = None, config: Config | None = None
    ) -> NDArrays:
        return self.pack_parameters(*self.select_parameters(model, initial_model))

    def pull_parameters(self, parameters: NDArrays, model: nn.Module, config: Config | None = None) -> None:
        arrays, names = self.unpack_parameters(parameters)
        inject_state(model, names, arrays)
