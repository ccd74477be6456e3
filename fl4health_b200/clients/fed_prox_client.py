"""FedProx client (parity: ``fl4health/clients/fed_prox_client.py:4-22``): the drift reference is the model received
at the start of the round."""

from __future__ import annotations

from fl4health_b200.clients.adaptive_drift_constraint_client import AdaptiveDriftConstraintClient


class FedProxClient(AdaptiveDriftConstraintClient):
    def update_before_train(self, current_server_round: int) -> None:
        self.drift_penalty_tensors = self.snapshot_drift_anchor()
        return super().update_before_train(current_server_round)
