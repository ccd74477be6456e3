"""SCAFFOLD server strategy.

Parity: ``fl4health/strategies/scaffold.py:28-424``: uniform mean of the packed ``[y_i..., delta_c_i...]``;
``x <- x + lr_s (ybar - x)``; ``c <- c + (|S|/N) * mean(delta_c)`` with ``|S|/N = fraction_fit``;
zero-initialised variates from a model.  Wire format ``weights ++ variates`` is unchanged.
"""

from __future__ import annotations

from collections.abc import Callable
from typing import Any

import numpy as np
import torch
from torch import nn

from fl4health_b200.client_managers.base_sampling_manager import BaseFractionSamplingManager
from fl4health_b200.common.typing import (
    FitIns,
    FitRes,
    MetricsAggregationFn,
    NDArrays,
    Parameters,
    Scalar,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
    to_tensor,
)
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithControlVariates
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.aggregate_utils import weighted_combine
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class Scaffold(BasicFedAvg):
    def __init__(
        self,
        *,
        fraction_fit: float = 1.0,
        fraction_evaluate: float = 1.0,
        min_available_clients: int = 2,
        evaluate_fn: Callable[[int, NDArrays, dict[str, Scalar]], tuple[float, dict[str, Scalar]] | None] | None = None,
        on_fit_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        on_evaluate_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        accept_failures: bool = True,
        initial_parameters: Parameters,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        weighted_eval_losses: bool = True,
        learning_rate: float = 1.0,
        initial_control_variates: Parameters | None = None,
        model: nn.Module | None = None,
    ) -> None:
        self.server_model_weights: NDArrays = parameters_to_ndarrays(initial_parameters)
        variates = self.initialize_control_variates(initial_control_variates, model)
        initial_parameters.tensors.extend(variates.tensors)
        initial_parameters.flat = None  # the packed list is no longer a pure arena view
        super().__init__(
            fraction_fit=fraction_fit, fraction_evaluate=fraction_evaluate, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn, weighted_aggregation=False,
            weighted_eval_losses=weighted_eval_losses,
        )
        self.learning_rate = learning_rate
        self.parameter_packer = ParameterPackerWithControlVariates(len(self.server_model_weights))

    def initialize_control_variates(self, initial_control_variates: Parameters | None, model: nn.Module | None) -> Parameters:
        if initial_control_variates is not None:
            self.server_control_variates: NDArrays = parameters_to_ndarrays(initial_control_variates)
            return initial_control_variates
        if model is not None:
            zeros = NDArrays([torch.zeros_like(p.data) for p in model.parameters() if p.requires_grad])
            self.server_control_variates = zeros
            return ndarrays_to_parameters(zeros)
        raise ValueError(
            "Both initial_control_variates and model are None. One must be defined in order to establish "
            "initial values for the control variates."
        )

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        # both halves of the packed payload (weights, variate updates) go through weighted_combine, which reduces SPMD
        # payloads with collectives: no need to broadcast every client's full payload first
        decoded = [arrays for _, arrays, _ in decode_and_pseudo_sort_results(results, materialize=False)]
        aggregated = self.aggregate(decoded)
        weights, variate_updates = self.parameter_packer.unpack_parameters(aggregated)
        self.server_model_weights = self.compute_updated_weights(weights)
        self.server_control_variates = self.compute_updated_control_variates(variate_updates)
        packed = self.parameter_packer.pack_parameters(self.server_model_weights, self.server_control_variates)
        return ndarrays_to_parameters(packed), self._aggregate_fit_metrics(server_round, results)

    def compute_parameter_delta(self, params_1: NDArrays, params_2: NDArrays) -> NDArrays:
        return NDArrays([_t(p1) - _t(p2, _t(p1)) for p1, p2 in zip(params_1, params_2)])

    def compute_updated_parameters(
        self, scaling_coefficient: float, original_params: NDArrays, parameter_updates: NDArrays
    ) -> NDArrays:
        out = NDArrays()
        for original, update in zip(original_params, parameter_updates):
            upd = _t(update)
            orig = _t(original, upd)
            if orig.is_floating_point():
                out.append(orig + scaling_coefficient * upd.to(orig.dtype))
            else:  # integer buffers (num_batches_tracked): carry the aggregated value through
                out.append((orig.double() + scaling_coefficient * upd.double()).to(orig.dtype))
        return out

    def aggregate(self, params: list[NDArrays]) -> NDArrays:
        """Uniform mean over clients of the packed payloads (weights part fused when arena-backed)."""
        k = len(params)
        split = self.parameter_packer.size_of_model_params
        unpacked = [self.parameter_packer.unpack_parameters(p) for p in params]
        weights = weighted_combine([u[0] for u in unpacked], [1.0 / k] * k)
        variates = weighted_combine([u[1] for u in unpacked], [1.0 / k] * k)
        assert len(weights) == split
        return self.parameter_packer.pack_parameters(weights, variates)

    def configure_fit_all(
        self, server_round: int, parameters: Parameters, client_manager: ClientManager
    ) -> list[tuple[ClientProxy, FitIns]]:
        """All clients participate (used by the warm-start round)."""
        assert isinstance(client_manager, BaseFractionSamplingManager)
        config = self.on_fit_config_fn(server_round) if self.on_fit_config_fn is not None else {"current_server_round": server_round}
        fit_ins = FitIns(parameters, config)
        return [(client, fit_ins) for client in client_manager.sample_all(self.min_available_clients)]

    def compute_updated_weights(self, weights: NDArrays) -> NDArrays:
        """x <- x + lr_s * (ybar - x)"""
        stepped = self._flat_server_step("x", self.server_model_weights, weights, self.learning_rate, towards=True)
        if stepped is not None:
            return stepped
        delta = self.compute_parameter_delta(weights, self.server_model_weights)
        return self.compute_updated_parameters(self.learning_rate, self.server_model_weights, delta)

    def compute_updated_control_variates(self, control_variates_update: NDArrays) -> NDArrays:
        """c <- c + (|S| / N) * mean(delta_c)"""
        stepped = self._flat_server_step("c", self.server_control_variates, control_variates_update, self.fraction_fit, towards=False)
        if stepped is not None:
            return stepped
        return self.compute_updated_parameters(self.fraction_fit, self.server_control_variates, control_variates_update)

    def _flat_server_step(self, which: str, current: NDArrays, incoming: NDArrays, scale: float, towards: bool) -> NDArrays | None:
        """The server update of one block as ONE launch over flat storage when the aggregate is arena-shaped:
        ``state.lerp_(incoming, scale)`` (``towards``: x <- x + scale (ybar - x)) or ``state.add_(incoming, alpha=scale)``
        (c <- c + scale * mean(delta_c)) instead of three element-wise kernels per tensor.  The strategy keeps the
        block in a flat buffer of its own (``state``); the returned list are per-tensor views of it."""
        flat, layout = getattr(incoming, "flat", None), getattr(incoming, "layout", None)
        if flat is None or layout is None or not flat.is_floating_point() or len(incoming) != len(current):
            return None
        store: dict[str, Any] = self.__dict__.setdefault("_flat_server_state", {})
        held = store.get(which)
        if held is None or held[0] is not layout or held[1].shape != flat.shape or held[1].device != flat.device:
            state = torch.zeros_like(flat)  # first arena-shaped aggregate: adopt the current values, tensor by tensor, once
            views = _views_of(layout, state)
            for view, value in zip(views, current):
                if view.is_floating_point():
                    view.copy_(_t(value, view))
            store[which] = held = (layout, state)
        state = held[1]
        if towards:
            state.lerp_(flat, float(scale))
        else:
            state.add_(flat, alpha=float(scale))
        out = _views_of(layout, state)
        integer_positions = [i for i, value in enumerate(incoming) if isinstance(value, torch.Tensor) and not value.is_floating_point()]
        if towards and scale == 1.0:
            out.int_flat = getattr(incoming, "int_flat", None)  # every counter is adopted as aggregated: so is their packed form
        for i in integer_positions:  # integer buffers (num_batches_tracked) are not in the float block: reference formula, per entry
            if towards and scale == 1.0:
                out[i] = incoming[i]  # x + 1 * (ybar - x)
                continue
            old, new = _t(current[i], incoming[i]).double(), incoming[i].double()
            out[i] = (old + scale * ((new - old) if towards else new)).to(incoming[i].dtype)
        return out


def _views_of(layout: Any, region: torch.Tensor) -> NDArrays:
    """A fresh list of per-tensor views of ``region`` (the layout's own list may be cached and shared)."""
    return NDArrays(layout.ndarrays(region=region), flat=region, layout=layout)


def _t(value: Any, like: torch.Tensor | None = None) -> torch.Tensor:
    tensor = to_tensor(value) if not isinstance(value, torch.Tensor) else value
    if like is not None and tensor.device != like.device:
        tensor = tensor.to(like.device)
    return tensor


class OpacusScaffold(Scaffold):
    """SCAFFOLD whose initial parameters / zero variates come from a DP-wrapped model (parity: ``scaffold.py:349-424``)."""

    def __init__(self, *, model: nn.Module, **kwargs: Any) -> None:
        from fl4health_b200.privacy.dp_engine import GradSampleModule

        assert isinstance(model, GradSampleModule), "Provided model must be a GradSampleModule"
        initial_parameters = ndarrays_to_parameters([v.detach().clone() for v in model.state_dict().values()])
        variates = ndarrays_to_parameters([torch.zeros_like(p.data) for p in model.parameters() if p.requires_grad])
        kwargs.pop("initial_parameters", None)
        super().__init__(initial_parameters=initial_parameters, initial_control_variates=variates, model=None, **kwargs)
