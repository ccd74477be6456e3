"""FedAvg over *ragged* layer sets: each client sends whichever layers it selected, keyed by name in the trailing slot
(parity: ``fl4health/strategies/fedavg_dynamic_layer.py:17-222``).  Per-name weighted / uniform means, on device."""

from __future__ import annotations

from collections import defaultdict
from typing import Any

import torch

from fl4health_b200.common.typing import FitRes, NDArray, NDArrays, Parameters, Scalar, ndarrays_to_parameters, to_tensor
from fl4health_b200.parameter_exchange.parameter_packer import ParameterPackerWithLayerNames
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class FedAvgDynamicLayer(BasicFedAvg):
    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.parameter_packer = ParameterPackerWithLayerNames()

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        decoded = [(arrays, n) for _, arrays, n in decode_and_pseudo_sort_results(results)]
        aggregated = self.aggregate(decoded)
        names = list(aggregated.keys())
        packed = self.parameter_packer.pack_parameters(NDArrays([aggregated[n] for n in names]), names)
        return ndarrays_to_parameters(packed), self._aggregate_fit_metrics(server_round, results)

    def aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, NDArray]:
        return self.weighted_aggregate(results) if self.weighted_aggregation else self.unweighted_aggregate(results)

    def _grouped(self, results: list[tuple[NDArrays, int]], weighted: bool) -> dict[str, NDArray]:
        sums: dict[str, torch.Tensor] = {}
        totals: defaultdict[str, float] = defaultdict(float)
        for packed_layers, num_examples in results:
            layers, names = self.parameter_packer.unpack_parameters(packed_layers)
            weight = float(num_examples) if weighted else 1.0
            for layer, name in zip(layers, names):
                tensor = to_tensor(layer).to(torch.float32) * weight
                sums[name] = tensor if name not in sums else sums[name] + tensor.to(sums[name].device)
                totals[name] += weight
        return {name: total / totals[name] for name, total in sums.items()}

    def weighted_aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, NDArray]:
        return self._grouped(results, weighted=True)

    def unweighted_aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, NDArray]:
        return self._grouped(results, weighted=False)
