"""Federated PCA (parity: ``fl4health/strategies/fedpca.py:18-270``): clients send ``(U_k, sigma_k)`` of their local
data; the server merges the subspaces either by one SVD of ``[U_1 S_1 | ... | U_K S_K]`` or by recursive QR merging.
Runs on whatever device the client arrays live on (torch.linalg)."""

from __future__ import annotations

from collections.abc import Callable
from logging import INFO
from typing import Any

import torch

from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import FitRes, MetricsAggregationFn, NDArrays, Parameters, Scalar, ndarrays_to_parameters, to_tensor
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results

EVALUATE_FN_TYPE = Callable[[int, NDArrays, dict[str, Scalar]], tuple[float, dict[str, Scalar]] | None] | None

MINIMUM_PCA_CLIENTS = 2
MINIMUM_PCA_ClIENTS = MINIMUM_PCA_CLIENTS  # the reference spells the constant with a lower-case "l"; both import


class FedPCA(BasicFedAvg):
    def __init__(
        self,
        *,
        fraction_fit: float = 1.0,
        fraction_evaluate: float = 1.0,
        min_fit_clients: int = 2,
        min_evaluate_clients: int = 2,
        min_available_clients: int = 2,
        evaluate_fn: Callable[[int, NDArrays, dict[str, Scalar]], tuple[float, dict[str, Scalar]] | None] | None = None,
        on_fit_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        on_evaluate_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        accept_failures: bool = True,
        initial_parameters: Parameters | None = None,
        fit_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        evaluate_metrics_aggregation_fn: MetricsAggregationFn | None = None,
        weighted_aggregation: bool = True,
        weighted_eval_losses: bool = True,
        svd_merging: bool = True,
    ) -> None:
        super().__init__(
            fraction_fit=fraction_fit, fraction_evaluate=fraction_evaluate, min_fit_clients=min_fit_clients,
            min_evaluate_clients=min_evaluate_clients, min_available_clients=min_available_clients,
            evaluate_fn=evaluate_fn, on_fit_config_fn=on_fit_config_fn, on_evaluate_config_fn=on_evaluate_config_fn,
            accept_failures=accept_failures, initial_parameters=initial_parameters,
            fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
            evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
            weighted_aggregation=weighted_aggregation, weighted_eval_losses=weighted_eval_losses,
        )
        self.svd_merging = svd_merging

    def aggregate_fit(self, server_round: int, results: list[tuple[ClientProxy, FitRes]], failures: list[Any]) -> tuple[Parameters | None, dict[str, Scalar]]:
        if not results or (not self.accept_failures and failures):
            return None, {}
        decoded = [arrays for _, arrays, _ in decode_and_pseudo_sort_results(results)]
        vectors = [to_tensor(a[0]).double() for a in decoded]
        values = [to_tensor(a[1]).double() for a in decoded]
        log(INFO, "Performing SVD-based merging." if self.svd_merging else "Performing QR-based merging.")
        merge = self.merge_subspaces_svd if self.svd_merging else self.merge_subspaces_qr
        merged_vectors, merged_values = merge(vectors, values)
        params = ndarrays_to_parameters(NDArrays([merged_vectors.float(), merged_values.float()]))
        return params, self._aggregate_fit_metrics(server_round, results)

    def merge_subspaces_svd(self, client_singular_vectors: list[torch.Tensor], client_singular_values: list[torch.Tensor]) -> tuple[torch.Tensor, torch.Tensor]:
        """SVD of the column-concatenation [U_1 diag(s_1) | ... | U_K diag(s_K)]."""
        stacked = torch.cat([u * s.unsqueeze(0) for u, s in zip(client_singular_vectors, client_singular_values)], dim=1)
        new_vectors, new_values, _ = torch.linalg.svd(stacked, full_matrices=True)
        return new_vectors, new_values

    def merge_subspaces_qr(self, client_singular_vectors: list[torch.Tensor], client_singular_values: list[torch.Tensor]) -> tuple[torch.Tensor, torch.Tensor]:
        assert len(client_singular_values) >= MINIMUM_PCA_CLIENTS
        u, s = client_singular_vectors[0], client_singular_values[0]
        for u_next, s_next in zip(client_singular_vectors[1:], client_singular_values[1:]):
            u, s = self.merge_two_subspaces_qr((u, torch.diag(s)), (u_next, torch.diag(s_next)))
        return u, s

    def merge_two_subspaces_qr(self, subspace1: tuple[torch.Tensor, torch.Tensor], subspace2: tuple[torch.Tensor, torch.Tensor]) -> tuple[torch.Tensor, torch.Tensor]:
        """Rank-revealing merge (Rehurek 2011): project U2 on U1, QR the residual, SVD the small core matrix."""
        (u1, s1), (u2, s2) = subspace1, subspace2
        z = u1.T @ u2
        q, r = torch.linalg.qr(u2 - u1 @ z)
        top = torch.cat([s1, z @ s2], dim=1)
        bottom = torch.cat([torch.zeros(r.shape[0], s1.shape[1], dtype=s1.dtype, device=s1.device), r @ s2], dim=1)
        u3, s_final, _ = torch.linalg.svd(torch.cat([top, bottom], dim=0), full_matrices=False)
        u_final = torch.cat([u1, q], dim=1) @ u3
        rank = min(u1.shape[0], u1.shape[1] + u2.shape[1])
        return u_final[:, :rank], s_final[:rank]
