"""Server-side checkpoint + state bundles.

Parity: ``fl4health/checkpointing/server_module.py:34-577``.  A server holds ``Parameters`` (not a module); to write a
model artifact the parameters are "hydrated" into a template ``nn.Module`` through an exchanger, after stripping any
packed side-information with the packer that matches the strategy.
"""

from __future__ import annotations

from collections.abc import Sequence
from logging import INFO
from typing import TYPE_CHECKING

from torch import nn

from fl4health_b200.checkpointing.checkpointer import TorchModuleCheckpointer
from fl4health_b200.checkpointing.opacus_checkpointer import OpacusCheckpointer
from fl4health_b200.checkpointing.state_checkpointer import NnUnetServerStateCheckpointer, ServerStateCheckpointer
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Parameters, Scalar, ndarrays_to_parameters, parameters_to_ndarrays
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import (
    ParameterPacker,
    ParameterPackerAdaptiveConstraint,
    ParameterPackerWithClippingBit,
    ParameterPackerWithControlVariates,
    ParameterPackerWithLayerNames,
    SparseCooParameterPacker,
)

if TYPE_CHECKING:
    from fl4health_b200.servers.base_server import FlServer

ModelCheckpointers = TorchModuleCheckpointer | Sequence[TorchModuleCheckpointer] | None


class BaseServerCheckpointAndStateModule:
    def __init__(
        self,
        model: nn.Module | None = None,
        parameter_exchanger: ParameterExchanger | None = None,
        model_checkpointers: ModelCheckpointers = None,
        state_checkpointer: ServerStateCheckpointer | None = None,
    ) -> None:
        self.model = model
        self.parameter_exchanger = parameter_exchanger
        self.model_checkpointers: list[TorchModuleCheckpointer] | None = (
            [model_checkpointers]
            if isinstance(model_checkpointers, TorchModuleCheckpointer)
            else (list(model_checkpointers) if model_checkpointers is not None else None)
        )
        self.state_checkpointer = state_checkpointer
        if self.model_checkpointers:
            self._validate_model_checkpointer_components()
        self._check_if_shared_checkpoint_names()

    def _validate_model_checkpointer_components(self) -> None:
        assert self.model is not None, "Checkpointer(s) defined but no model is defined to hydrate."
        assert self.parameter_exchanger is not None, "Checkpointer(s) defined but no parameter_exchanger to hydrate."

    def _check_if_shared_checkpoint_names(self) -> None:
        paths = [c.checkpoint_path for c in (self.model_checkpointers or [])]
        if len(set(paths)) != len(paths):
            listing = "\n".join(paths)
            raise ValueError(
                "The paths of all of your checkpointers should be unique otherwise overwrites are possible and data "
                f"will be lost. The current paths are:\n{listing}"
            )

    def maybe_checkpoint(self, server_parameters: Parameters, loss: float, metrics: dict[str, Scalar]) -> None:
        if not self.model_checkpointers:
            log(INFO, "No model checkpointers specified. Skipping any checkpointing.")
            return
        assert self.model is not None
        self._hydrate_model_for_checkpointing(server_parameters)
        for checkpointer in self.model_checkpointers:
            checkpointer.maybe_checkpoint(self.model, loss, metrics)

    def _hydrate_model_for_checkpointing(self, server_parameters: Parameters) -> None:
        assert self.model is not None, "Hydrate model for checkpoint called but self.model is None"
        assert self.parameter_exchanger is not None, "Hydrate model called but self.parameter_exchanger is None"
        self.parameter_exchanger.pull_parameters(parameters_to_ndarrays(server_parameters), self.model)

    def save_state(self, server: FlServer, server_parameters: Parameters) -> None:
        if self.state_checkpointer is None:
            raise ValueError("Attempting to save state but no state checkpointer is specified")
        self._hydrate_model_for_checkpointing(server_parameters)
        assert self.model is not None
        self.state_checkpointer.save_server_state(server, self.model)

    def maybe_load_state(self, server: FlServer) -> Parameters | None:
        if self.state_checkpointer is None:
            raise ValueError("Attempting to load state, but no state checkpointer is specified")
        assert self.model is not None, "Attempting to load state but self.model is None"
        server_model = self.state_checkpointer.maybe_load_server_state(server, self.model)
        if server_model is None:
            return None
        assert self.parameter_exchanger is not None
        return ndarrays_to_parameters(self.parameter_exchanger.push_parameters(server_model))


class PackingServerCheckpointAndAndStateModule(BaseServerCheckpointAndStateModule):
    def __init__(
        self,
        model: nn.Module | None = None,
        parameter_exchanger: FullParameterExchangerWithPacking | None = None,
        model_checkpointers: ModelCheckpointers = None,
        state_checkpointer: ServerStateCheckpointer | None = None,
    ) -> None:
        if parameter_exchanger is not None:
            assert isinstance(parameter_exchanger, FullParameterExchangerWithPacking), (
                "Parameter exchanger must be of based type FullParameterExchangerWithPacking"
            )
        super().__init__(model, parameter_exchanger, model_checkpointers, state_checkpointer)

    def _hydrate_model_for_checkpointing(self, server_parameters: Parameters) -> None:
        assert self.model is not None and isinstance(self.parameter_exchanger, FullParameterExchangerWithPacking)
        weights, _ = self.parameter_exchanger.unpack_parameters(parameters_to_ndarrays(server_parameters))
        self.parameter_exchanger.pull_parameters(weights, self.model)


def _packing_module(packer_factory):  # noqa: ANN001, ANN202
    """Build a module class whose exchanger is ``FullParameterExchangerWithPacking(packer_factory(model))``."""

    class _Module(PackingServerCheckpointAndAndStateModule):
        def __init__(
            self,
            model: nn.Module | None = None,
            model_checkpointers: ModelCheckpointers = None,
            state_checkpointer: ServerStateCheckpointer | None = None,
        ) -> None:
            exchanger = None
            if model is not None:
                packer: ParameterPacker = packer_factory(model)
                exchanger = FullParameterExchangerWithPacking(packer)
            super().__init__(model, exchanger, model_checkpointers, state_checkpointer)

    return _Module


class ScaffoldServerCheckpointAndStateModule(
    _packing_module(lambda model: ParameterPackerWithControlVariates(len(model.state_dict())))  # type: ignore[misc]
):
    """Strips the control variates appended after the ``len(state_dict)`` weight arrays."""


class AdaptiveConstraintServerCheckpointAndStateModule(
    _packing_module(lambda model: ParameterPackerAdaptiveConstraint())  # type: ignore[misc]
):
    """Strips the trailing drift-penalty weight (FedProx / Ditto / MR-MTL)."""


class ClippingBitServerCheckpointAndStateModule(
    _packing_module(lambda model: ParameterPackerWithClippingBit())  # type: ignore[misc]
):
    """Strips the trailing clipping bound (client-level DP)."""


class LayerNamesServerCheckpointAndStateModule(
    _packing_module(lambda model: ParameterPackerWithLayerNames())  # type: ignore[misc]
):
    """Strips the trailing layer-name array (dynamic layer exchange)."""

    def _hydrate_model_for_checkpointing(self, server_parameters: Parameters) -> None:
        from fl4health_b200.parameter_exchange._state import inject_state

        assert self.model is not None and isinstance(self.parameter_exchanger, FullParameterExchangerWithPacking)
        weights, names = self.parameter_exchanger.unpack_parameters(parameters_to_ndarrays(server_parameters))
        inject_state(self.model, names, weights)


class SparseCooServerCheckpointAndStateModule(
    _packing_module(lambda model: SparseCooParameterPacker())  # type: ignore[misc]
):
    """Densifies COO payloads before hydrating."""

    def _hydrate_model_for_checkpointing(self, server_parameters: Parameters) -> None:
        from fl4health_b200.parameter_exchange.sparse_coo_parameter_exchanger import SparseCooParameterExchanger

        assert self.model is not None
        SparseCooParameterExchanger(1.0, lambda m, i: {}).pull_parameters(
            parameters_to_ndarrays(server_parameters), self.model
        )


class OpacusServerCheckpointAndStateModule(BaseServerCheckpointAndStateModule):
    def __init__(
        self,
        model: nn.Module | None = None,
        parameter_exchanger: ParameterExchanger | None = None,
        model_checkpointers: ModelCheckpointers = None,
        state_checkpointer: ServerStateCheckpointer | None = None,
    ) -> None:
        super().__init__(model, parameter_exchanger, model_checkpointers, state_checkpointer)
        for checkpointer in self.model_checkpointers or []:
            assert isinstance(checkpointer, OpacusCheckpointer), "Provided checkpointers must be OpacusCheckpointers"


class NnUnetServerCheckpointAndStateModule(BaseServerCheckpointAndStateModule):
    """The model may be attached later (architecture is negotiated with clients before round 1)."""

    def __init__(
        self,
        model: nn.Module | None = None,
        parameter_exchanger: ParameterExchanger | None = None,
        model_checkpointers: ModelCheckpointers = None,
        state_checkpointer: NnUnetServerStateCheckpointer | None = None,
    ) -> None:
        self.model = model
        self.parameter_exchanger = parameter_exchanger or FullParameterExchanger()
        self.model_checkpointers = (
            [model_checkpointers]
            if isinstance(model_checkpointers, TorchModuleCheckpointer)
            else (list(model_checkpointers) if model_checkpointers is not None else None)
        )
        self.state_checkpointer = state_checkpointer
        self._check_if_shared_checkpoint_names()


class DpScaffoldServerCheckpointAndStateModule(ScaffoldServerCheckpointAndStateModule):
    def __init__(
        self,
        model: nn.Module | None = None,
        model_checkpointers: ModelCheckpointers = None,
        state_checkpointer: ServerStateCheckpointer | None = None,
    ) -> None:
        super().__init__(model, model_checkpointers, state_checkpointer)
        for checkpointer in self.model_checkpointers or []:
            assert isinstance(checkpointer, OpacusCheckpointer), "Provided checkpointers must be OpacusCheckpointers"
