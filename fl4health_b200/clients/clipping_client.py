"""Client-level DP client (parity: ``fl4health/clients/clipping_client.py:22-188``): uploads the *clipped weight delta*
``clip(w_new - w_start, C)`` plus a clipping bit (1 if the update was already within the bound; only meaningful with
adaptive clipping).  The reference computes the norm layer by layer in NumPy on the CPU; here the norm of the whole
delta is one device reduction and the scaling one fused pass (arena) or a handful of device ops (no arena)."""

from __future__ import annotations

from logging import INFO
from typing import Any

import torch

from fl4health_b200.clients.basic_client import BasicClient
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, NDArrays, to_tensor
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithClippingBit
from fl4health_b200.utils.config import narrow_dict_type


class NumpyClippingClient(BasicClient):
    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.parameter_exchanger: FullParameterExchangerWithPacking[float]
        self.clipping_bound: float | None = None
        self.adaptive_clipping: bool | None = None

    def calculate_parameters_norm(self, parameters: NDArrays) -> float:
        """Frobenius norm over ALL layers (one device reduction + one read-back)."""
        squares = [to_tensor(layer).double().pow(2).sum() for layer in parameters]
        return float(torch.stack(squares).sum().sqrt().item())

    def clip_parameters(self, parameters: NDArrays) -> tuple[NDArrays, float]:
        assert self.clipping_bound is not None and self.adaptive_clipping is not None
        norm = self.calculate_parameters_norm(parameters)
        log(INFO, f"Update norm: {norm}, Clipping Bound: {self.clipping_bound}")
        if norm <= self.clipping_bound:
            return parameters, (1.0 if self.adaptive_clipping else 0.0)
        scale = min(1.0, self.clipping_bound / norm)
        return NDArrays([to_tensor(layer) * scale for layer in parameters]), 0.0

    def compute_weight_update_and_clip(self, parameters: NDArrays) -> tuple[NDArrays, float]:
        assert self.initial_weights is not None and len(parameters) == len(self.initial_weights)
        update = NDArrays([to_tensor(new) - to_tensor(old, to_tensor(new).device) for old, new in zip(self.initial_weights, parameters)])
        return self.clip_parameters(update)

    def get_parameters(self, config: Config) -> NDArrays:
        current_server_round = int(config.get("current_server_round", 0))
        if not self.initialized or current_server_round == 0:
            return self.setup_client_and_return_all_model_parameters(config)
        assert self.model is not None and self.parameter_exchanger is not None
        model_weights = self.parameter_exchanger.push_parameters(self.model, config=config)
        clipped_update, clipping_bit = self.compute_weight_update_and_clip(model_weights)
        return self.parameter_exchanger.pack_parameters(clipped_update, clipping_bit)

    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        assert self.model is not None and self.parameter_exchanger is not None
        server_model_parameters, clipping_bound = self.parameter_exchanger.unpack_parameters(parameters)
        self.clipping_bound = clipping_bound
        current_server_round = narrow_dict_type(config, "current_server_round", int)
        if current_server_round == 1 and fitting_round:
            self.initialize_all_model_weights(server_model_parameters, config)
        else:
            self.parameter_exchanger.pull_parameters(server_model_parameters, self.model, config)
        # snapshot of the round-start weights (detached from the live model: training mutates it in place)
        self.initial_weights = NDArrays([t.detach().clone() for t in self.parameter_exchanger.push_parameters(self.model, config=config)])

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchangerWithPacking(ParameterPackerWithClippingBit())

    def setup_client(self, config: Config) -> None:
        self.adaptive_clipping = narrow_dict_type(config, "adaptive_clipping", bool)
        super().setup_client(config)
