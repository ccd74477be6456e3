"""FedAvg over per-tensor COO payloads (parity: ``fl4health/strategies/fedavg_sparse_coo_tensor.py:18-316``).
Each client's (values, indices, shape) triple is scattered straight into a dense device accumulator with
``index_put_(accumulate=True)`` — no ``torch.sparse_coo_tensor(...).to_dense()`` temporary per client per tensor."""

from __future__ import annotations

from collections import defaultdict
from typing import Any

import numpy as np
import torch

from fl4health_b200.common.typing import FitRes, NDArrays, Parameters, Scalar, ndarrays_to_parameters, to_tensor
from fl4health_b200.parameter_exchange.parameter_packer import SparseCooParameterPacker
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results


class FedAvgSparseCooTensor(BasicFedAvg):
    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.parameter_packer = SparseCooParameterPacker()

    def aggregate_fit(
        self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]
    ) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        decoded = [(arrays, n) for _, arrays, n in decode_and_pseudo_sort_results(results)]
        aggregated = self.aggregate(decoded)
        names, values, indices, shapes = [], NDArrays(), NDArrays(), NDArrays()
        for name, dense in aggregated.items():
            vals, idx, shape = self.parameter_packer.extract_coo_info_from_dense(dense)
            names.append(name)
            values.append(vals)
            indices.append(idx)
            shapes.append(shape)
        packed = self.parameter_packer.pack_parameters(values, (indices, shapes, names))
        return ndarrays_to_parameters(packed), self._aggregate_fit_metrics(server_round, results)

    def aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, torch.Tensor]:
        return self.weighted_aggregate(results) if self.weighted_aggregation else self.unweighted_aggregate(results)

    def _accumulate(self, results: list[tuple[NDArrays, int]], weighted: bool) -> dict[str, torch.Tensor]:
        dense: dict[str, torch.Tensor] = {}
        totals: defaultdict[str, float] = defaultdict(float)
        for packed, num_examples in results:
            values, (indices, shapes, names) = self.parameter_packer.unpack_parameters(packed)
            assert len(values) == len(indices) == len(shapes) == len(names) and len(names) > 0
            weight = float(num_examples) if weighted else 1.0
            for vals, idx, shape, name in zip(values, indices, shapes, names):
                v = to_tensor(vals).to(torch.float32)
                i = to_tensor(idx, v.device).long()
                if name not in dense:
                    dense[name] = torch.zeros(tuple(int(s) for s in np.asarray(shape).tolist()), dtype=torch.float32, device=v.device)
                if i.numel() > 0:
                    dense[name].index_put_(tuple(i.t()), v.to(dense[name].device) * weight, accumulate=True)
                totals[name] += weight
        return {name: acc / totals[name] for name, acc in dense.items()}

    def weighted_aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, torch.Tensor]:
        return self._accumulate(results, weighted=True)

    def unweighted_aggregate(self, results: list[tuple[NDArrays, int]]) -> dict[str, torch.Tensor]:
        return self._accumulate(results, weighted=False)
