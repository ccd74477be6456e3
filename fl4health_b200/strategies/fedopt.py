"""Server-optimizer strategies: FedOpt / FedAdam / FedAdagrad / FedYogi.

The reference imports these from Flower (``examples/fedopt_example/server.py:138-151``):
``delta = avg(w_k) - w``; ``m <- b1 m + (1-b1) delta``; ``v`` per variant; ``w <- w + eta m / (sqrt(v) + tau)``
(defaults eta=1e-1, eta_l=1e-1, b1=0.9, b2=0.99, tau=1e-9).  Here the mean, the moment updates and the weight update
are ONE pass (``ops.flat.weighted_sum`` with an epilogue) when clients are arena-backed; the per-layer fallback uses
device tensor ops.  Server moments live in flat buffers with the arena's offsets.
"""

from __future__ import annotations

from typing import Any

import torch

from fl4health_b200.common.typing import (
    FitRes,
    NDArrays,
    Parameters,
    Scalar,
    ndarrays_to_parameters,
    parameters_to_ndarrays,
    to_tensor,
)
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.aggregate_utils import _common_flat, _spmd_weighted_combine, aggregate_results
from fl4health_b200.strategies.fedavg import FedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class FedOpt(FedAvg):
    _mode = flat_ops.EPI_NONE

    def __init__(
        self,
        *,
        initial_parameters: Parameters,
        eta: float = 1e-1,
        eta_l: float = 1e-1,
        beta_1: float = 0.0,
        beta_2: float = 0.0,
        tau: float = 1e-9,
        **kwargs: Any,
    ) -> None:
        super().__init__(initial_parameters=initial_parameters, **kwargs)
        self.current_weights: NDArrays = parameters_to_ndarrays(initial_parameters)
        self.eta, self.eta_l, self.beta_1, self.beta_2, self.tau = eta, eta_l, beta_1, beta_2, tau
        self.m_t: list[torch.Tensor] | None = None
        self.v_t: list[torch.Tensor] | None = None
        self._flat_m: torch.Tensor | None = None
        self._flat_v: torch.Tensor | None = None

    # per-layer second-moment rule, overridden by the variants
    def _second_moment(self, v: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        decoded = decode_and_pseudo_sort_results(results, materialize=False)
        client_arrays = [arrays for _, arrays, _ in decoded]
        counts = [n for _, _, n in decoded]
        metrics = self._aggregate_fit_metrics(server_round, results)

        current_flat = getattr(self.current_weights, "flat", None)
        spmd = any(getattr(a, "ctx", None) is not None for a in client_arrays)
        if spmd and current_flat is not None and self._mode != flat_ops.EPI_NONE and all(
            a.spec.is_arena and a.spec.flat_numel == current_flat.numel() for a in client_arrays
        ):
            if self._flat_m is None:
                self._flat_m = torch.zeros_like(current_flat)
                self._flat_v = torch.zeros_like(current_flat)
            total = float(sum(counts))
            epilogue = dict(mode=self._mode, current=current_flat, m=self._flat_m, v=self._flat_v, eta=self.eta,
                            beta1=self.beta_1, beta2=self.beta_2, tau=self.tau)
            new_weights = _spmd_weighted_combine(client_arrays, [n / total for n in counts], epilogue=epilogue)
            self.current_weights = new_weights
            return ndarrays_to_parameters(new_weights), metrics
        flats = _common_flat(client_arrays) if not spmd else None
        if flats is not None and current_flat is not None and current_flat.numel() == flats[0].numel() and self._mode != flat_ops.EPI_NONE:
            layout = client_arrays[0].layout
            if self._flat_m is None:
                self._flat_m = torch.zeros_like(current_flat)
                self._flat_v = torch.zeros_like(current_flat)
            total = float(sum(counts))
            new_flat = torch.empty_like(current_flat)
            flat_ops.weighted_sum(
                new_flat, flats, [n / total for n in counts], mode=self._mode, current=current_flat, m=self._flat_m,
                v=self._flat_v, eta=self.eta, beta1=self.beta_1, beta2=self.beta_2, tau=self.tau,
            )
            new_weights = layout.ndarrays(region=new_flat)
            mean_for_ints = aggregate_results([(a, n) for a, n in zip(client_arrays, counts)], weighted=True) if layout.int_state else None
            if mean_for_ints is not None:
                for idx, key in enumerate(layout.state_keys):
                    if key in layout.int_state:
                        new_weights[idx] = mean_for_ints[idx]
            self.current_weights = new_weights
            return ndarrays_to_parameters(new_weights), metrics

        mean = aggregate_results([(a, n) for a, n in zip(client_arrays, counts)], weighted=True)
        device = next((t.device for t in mean if isinstance(t, torch.Tensor)), None)
        current = [to_tensor(w, device) for w in self.current_weights]
        mean_t = [to_tensor(w, device) for w in mean]
        if self.m_t is None:
            self.m_t = [torch.zeros_like(w, dtype=torch.float32) for w in current]
            self.v_t = [torch.zeros_like(w, dtype=torch.float32) for w in current]
        assert self.v_t is not None
        new_weights = NDArrays()
        for idx, (w, avg) in enumerate(zip(current, mean_t)):
            if not w.is_floating_point():
                new_weights.append(avg.to(w.dtype))
                continue
            delta = avg.to(torch.float32) - w.to(torch.float32)
            self.m_t[idx] = self.beta_1 * self.m_t[idx] + (1.0 - self.beta_1) * delta
            self.v_t[idx] = self._second_moment(self.v_t[idx], delta)
            new_weights.append((w.to(torch.float32) + self.eta * self.m_t[idx] / (self.v_t[idx].sqrt() + self.tau)).to(w.dtype))
        self.current_weights = new_weights
        return ndarrays_to_parameters(new_weights), metrics


class FedAdam(FedOpt):
    _mode = flat_ops.EPI_FEDADAM

    def __init__(self, *, initial_parameters: Parameters, eta: float = 1e-1, eta_l: float = 1e-1,
                 beta_1: float = 0.9, beta_2: float = 0.99, tau: float = 1e-9, **kwargs: Any) -> None:
        super().__init__(initial_parameters=initial_parameters, eta=eta, eta_l=eta_l, beta_1=beta_1, beta_2=beta_2,
                         tau=tau, **kwargs)

    def _second_moment(self, v: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
        return self.beta_2 * v + (1.0 - self.beta_2) * delta * delta


class FedAdagrad(FedOpt):
    _mode = flat_ops.EPI_FEDADAGRAD

    def __init__(self, *, initial_parameters: Parameters, eta: float = 1e-1, eta_l: float = 1e-1,
                 tau: float = 1e-9, **kwargs: Any) -> None:
        super().__init__(initial_parameters=initial_parameters, eta=eta, eta_l=eta_l, beta_1=0.0, beta_2=0.0, tau=tau,
                         **kwargs)

    def _second_moment(self, v: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
        return v + delta * delta


class FedYogi(FedOpt):
    _mode = flat_ops.EPI_FEDYOGI

    def __init__(self, *, initial_parameters: Parameters, eta: float = 1e-2, eta_l: float = 0.0316,
                 beta_1: float = 0.9, beta_2: float = 0.99, tau: float = 1e-3, **kwargs: Any) -> None:
        super().__init__(initial_parameters=initial_parameters, eta=eta, eta_l=eta_l, beta_1=beta_1, beta_2=beta_2,
                         tau=tau, **kwargs)

    def _second_moment(self, v: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
        d2 = delta * delta
        return v - (1.0 - self.beta_2) * d2 * torch.sign(v - d2)
