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
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.parallel.arena import arena_of
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.parameter_exchange.parameter_exchanger_base import ParameterExchanger
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithClippingBit
from fl4health_b200.utils.config import narrow_dict_type


class NumpyClippingClient(BasicClient):
    """Constructor arguments are ``BasicClient``'s.  With an arena-backed model the whole delta pipeline stays on the
    device and in three launches: ``||w - w_start||^2`` (one reduction over the flat arena against its round-start
    snapshot), ``delta = w - w_start``, and ``delta *= min(1, C / ||delta||)`` which also writes the clipping bit — no
    per-layer loop and no host read-back of the norm (SURVEY hot-op L12)."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.parameter_exchanger: FullParameterExchangerWithPacking[float]
        self.clipping_bound: float | None = None
        self.adaptive_clipping: bool | None = None

    def setup_client(self, config: Config) -> None:
        self.adaptive_clipping = narrow_dict_type(config, "adaptive_clipping", bool)
        super().setup_client(config)

    def get_parameter_exchanger(self, config: Config) -> ParameterExchanger:
        return FullParameterExchangerWithPacking(ParameterPackerWithClippingBit())

    # ------------------------------------------------------------------------------------------ exchange
    def set_parameters(self, parameters: NDArrays, config: Config, fitting_round: bool) -> None:
        weights, self.clipping_bound = self.parameter_exchanger.unpack_parameters(parameters)
        BasicClient.set_parameters(self, weights, config, fitting_round)
        self._remember_round_start(config)

    def _remember_round_start(self, config: Config) -> None:
        """Snapshot of the weights local training starts from (training mutates the live model in place)."""
        arena = arena_of(self.model)
        if arena is not None:
            arena.companion("round_start").copy_(arena.flat)  # one flat copy
            ints = [t.detach().clone() for t in arena.int_state.values()]
            self.initial_weights = NDArrays(arena.ndarrays(region=arena.regions["round_start"]))
            for slot, value in zip(arena._int_positions(), ints):
                self.initial_weights[slot] = value
            return
        live = self.parameter_exchanger.push_parameters(self.model, config=config)
        self.initial_weights = NDArrays([t.detach().clone() for t in live])

    def get_parameters(self, config: Config) -> NDArrays:
        if not self.initialized or int(config.get("current_server_round", 0)) == 0:
            return self.setup_client_and_return_all_model_parameters(config)
        weights = self.parameter_exchanger.push_parameters(self.model, config=config)
        clipped_delta, clipping_bit = self.compute_weight_update_and_clip(weights)
        return self.parameter_exchanger.pack_parameters(clipped_delta, clipping_bit)

    # ------------------------------------------------------------------------------------------ clipping
    def calculate_parameters_norm(self, parameters: NDArrays) -> float:
        """Frobenius norm over ALL layers (one device reduction + one read-back)."""
        squares = torch.stack([to_tensor(layer).double().square().sum() for layer in parameters])
        return float(squares.sum().sqrt().item())

    def clip_parameters(self, parameters: NDArrays) -> tuple[NDArrays, float]:
        assert self.clipping_bound is not None and self.adaptive_clipping is not None
        norm = self.calculate_parameters_norm(parameters)
        log(INFO, f"Update norm: {norm}, Clipping Bound: {self.clipping_bound}")
        within = norm <= self.clipping_bound
        if within:
            return parameters, float(self.adaptive_clipping)
        shrink = self.clipping_bound / norm
        return NDArrays([to_tensor(layer) * shrink for layer in parameters]), 0.0

    def _flat_update_and_clip(self, arena: Any) -> tuple[NDArrays, Any]:
        assert self.clipping_bound is not None and self.adaptive_clipping is not None
        start, delta = arena.regions["round_start"], arena.companion("update")
        squared_norm = flat_ops.sq_diff_sum(arena.flat, start)
        counters = list(arena.int_state.values())
        old_counters = [self.initial_weights[slot] for slot in arena._int_positions()]  # type: ignore[index]
        counter_deltas = [new - to_tensor(old, new.device) for new, old in zip(counters, old_counters)]
        for step in counter_deltas:  # integer buffers are part of the update the reference clips (a handful of scalars)
            squared_norm += step.double().square().sum().float()
        torch.sub(arena.flat, start, out=delta)
        bit = torch.zeros(1, dtype=torch.float32, device=delta.device)
        flat_ops.clip_scale_(delta, squared_norm, float(self.clipping_bound), bit)
        clipped = NDArrays(arena.ndarrays(region=delta))
        scale = torch.clamp(self.clipping_bound / (squared_norm.sqrt() + 1e-12), max=1.0)
        for slot, step in zip(arena._int_positions(), counter_deltas):
            clipped[slot] = step * scale
        clipped.flat = None  # a packed (weights ++ bit) payload is not arena-shaped
        return clipped, (bit if self.adaptive_clipping else torch.zeros_like(bit))

    def compute_weight_update_and_clip(self, parameters: NDArrays) -> tuple[NDArrays, float]:
        assert self.initial_weights is not None and len(parameters) == len(self.initial_weights)
        model = getattr(self, "model", None)
        arena = arena_of(model) if model is not None else None
        if arena is not None and "round_start" in arena.regions and getattr(parameters, "flat", None) is not None \
                and parameters.flat.data_ptr() == arena.flat.data_ptr():
            return self._flat_update_and_clip(arena)
        delta = NDArrays([to_tensor(new) - to_tensor(old, to_tensor(new).device)
                          for old, new in zip(self.initial_weights, parameters)])
        return self.clip_parameters(delta)
