"""FedPM server strategy (Isik et al. 2023; parity: ``fl4health/strategies/fedpm.py:12-162``).

Clients upload Bernoulli-sampled binary masks.  Aggregation is the uniform mean, or -- ``bayesian_aggregation`` -- a
Beta posterior per score: ``alpha += sum M``, ``beta += K - sum M``, ``theta = (alpha - 1) / (alpha + beta - 2)``.

Two execution forms of the same vote:

* every client's masks are on this process (simulation, materialised payloads): one vote kernel per tensor over uint8
  masks (``ops.flat.fedpm_vote``);
* one client per rank (SPMD): the masks never travel as bytes.  Each rank packs its own masks to 1 bit per score
  (``ops.flat.pack_mask_bits``, a warp ballot per 32 scores), ONE all-gather moves ``K * n / 32`` words, and a single
  kernel counts the bits and updates the posterior for every layer at once (``ops.flat.fedpm_vote_packed``): 1/8 of the
  uint8 traffic, 1/32 of fp32, and no per-client materialising broadcasts.
"""

from __future__ import annotations

import math
from collections import defaultdict
from typing import Any

import numpy as np
import torch

from fl4health_b200.common.typing import FitRes, NDArray, NDArrays, Parameters, Scalar, ndarrays_to_parameters, to_tensor
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.fedavg_dynamic_layer import FedAvgDynamicLayer
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class FedPm(FedAvgDynamicLayer):
    def __init__(self, *, bayesian_aggregation: bool = True, **kwargs: Any) -> None:
        kwargs["weighted_aggregation"] = False
        super().__init__(**kwargs)
        self.beta_parameters: dict[str, tuple[torch.Tensor, torch.Tensor]] = {}
        self.bayesian_aggregation = bayesian_aggregation
        self._flat_priors: tuple[tuple, torch.Tensor, torch.Tensor] | None = None  # (layout key, alpha, beta) of the packed path
        self.last_vote_path = "none"  # "packed-bits" | "per-tensor": which form the latest aggregate took (tests, tracing)

    # ------------------------------------------------------------------------------------------ cross-rank bit vote
    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        payloads = [arrays for _, arrays, _ in decode_and_pseudo_sort_results(results, materialize=False)]
        plan = self._cross_rank_plan(payloads)
        if plan is None:
            return super().aggregate_fit(server_round, results, failures)
        ctx, mine, names, shapes = plan
        aggregated = self._vote_on_packed_bits(ctx, mine, names, shapes)
        packed = self.parameter_packer.pack_parameters(NDArrays([aggregated[name] for name in names]), names)
        return ndarrays_to_parameters(packed), self._aggregate_fit_metrics(server_round, results)

    @staticmethod
    def _cross_rank_plan(payloads: list[NDArrays]) -> tuple[Any, NDArrays, list[str], list[tuple[int, ...]]] | None:
        """The packed vote applies when the federation is one client per rank, every rank took part, and all clients
        report the same score tensors (FedPM clients always send every masked layer).  Everything is decided from the
        payload *specs*, which every rank holds for every client -- no tensor is touched."""
        ctx = getattr(payloads[0], "ctx", None)
        if ctx is None or ctx.world_size == 1 or len(payloads) != ctx.world_size:
            return None
        if sorted(getattr(p, "rank", -1) for p in payloads) != list(range(ctx.world_size)):
            return None
        signatures = []
        for payload in payloads:
            entries = payload.spec.entries
            names = entries[-1][2]  # trailing slot of the layer-name packer: the names, carried by value
            if names is None or any(inline is not None for _, _, inline in entries[:-1]):
                return None
            signatures.append(([str(n) for n in np.asarray(names).reshape(-1)], [shape for shape, _, _ in entries[:-1]]))
        if any(signature != signatures[0] for signature in signatures[1:]):
            return None
        mine = next(p for p in payloads if p.rank == ctx.rank)
        return ctx, mine, signatures[0][0], signatures[0][1]

    def _priors_for(self, names: list[str], shapes: list[tuple[int, ...]], device: torch.device) -> tuple[torch.Tensor, torch.Tensor]:
        """alpha / beta of all score tensors as two flat buffers (``beta_parameters[name]`` are views into them), so the
        whole model is one vote launch.  Evidence accumulated by the per-tensor path is carried over."""
        key = (tuple(names), tuple(shapes), str(device))
        if self._flat_priors is not None and self._flat_priors[0] == key:
            return self._flat_priors[1], self._flat_priors[2]
        sizes = [math.prod(shape) for shape in shapes]
        alpha, beta = (torch.ones(sum(sizes), dtype=torch.float32, device=device) for _ in range(2))
        offset = 0
        for name, shape, size in zip(names, shapes, sizes):
            views = alpha[offset:offset + size].view(shape), beta[offset:offset + size].view(shape)
            if name in self.beta_parameters:
                for view, old in zip(views, self.beta_parameters[name]):
                    view.copy_(old.to(device))
            self.beta_parameters[name] = views
            offset += size
        self._flat_priors = (key, alpha, beta)
        return alpha, beta

    def _vote_on_packed_bits(self, ctx: Any, mine: NDArrays, names: list[str], shapes: list[tuple[int, ...]]) -> dict[str, NDArray]:
        masks, _ = self.parameter_packer.unpack_parameters(mine)
        local = torch.cat([to_tensor(mask).to(ctx.device).reshape(-1) for mask in masks])
        total = local.numel()
        words = ctx.all_gather_rows(flat_ops.pack_mask_bits(local))  # [K, ceil(n / 32)] int32: the round's whole uplink
        alpha = beta = None
        if self.bayesian_aggregation:
            alpha, beta = self._priors_for(names, shapes, ctx.device)
        theta = flat_ops.fedpm_vote_packed(words, total, alpha, beta, bayesian=self.bayesian_aggregation)
        self.last_vote_path = "packed-bits"
        out: dict[str, NDArray] = {}
        offset = 0
        for name, shape in zip(names, shapes):
            size = math.prod(shape)
            out[name] = theta[offset:offset + size].view(shape)
            offset += size
        return out

    # ------------------------------------------------------------------------------------------ per-tensor vote
    def aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, NDArray]:
        self.last_vote_path = "per-tensor"
        if not self.bayesian_aggregation:
            return super().aggregate(results)
        return self.aggregate_bayesian(results)

    def aggregate_bayesian(self, results: list[tuple[NDArrays, int]]) -> dict[str, NDArray]:
        names_to_masks: defaultdict[str, list[torch.Tensor]] = defaultdict(list)
        for packed_layers, _ in results:
            layers, names = self.parameter_packer.unpack_parameters(packed_layers)
            for layer, name in zip(layers, names):
                mask = to_tensor(layer)
                names_to_masks[name].append(mask)
                if name not in self.beta_parameters:
                    ones = torch.ones(mask.shape, dtype=torch.float32, device=mask.device)
                    self.beta_parameters[name] = (ones, ones.clone())
        out: dict[str, NDArray] = {}
        for name, (alpha, beta) in self.beta_parameters.items():
            masks = names_to_masks.get(name)
            if not masks:
                continue
            flat_masks = [m.to(device=alpha.device, dtype=torch.uint8).reshape(-1).contiguous() for m in masks]
            theta = flat_ops.fedpm_vote(flat_masks, alpha.reshape(-1), beta.reshape(-1), bayesian=True)
            out[name] = theta.reshape(alpha.shape)
        return out

    def reset_beta_priors(self) -> None:
        """Forget accumulated evidence (called every ``reset_frequency`` rounds by ``FedPmServer``).  In place, so the
        per-name views of the flat prior buffers stay views."""
        for alpha, beta in self.beta_parameters.values():
            alpha.fill_(1.0)
            beta.fill_(1.0)
