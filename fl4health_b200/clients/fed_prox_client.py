"""FedProx client (parity: ``fl4health/clients/fed_prox_client.py:4-22``): the drift reference is the model received
at the start of the round — in the engine the pull kernel writes it (``ParameterArena._fused_pull``)."""

from __future__ import annotations

from fl4health_b200.clients.adaptive_drift_constraint_client import AdaptiveDriftConstraintClient


class FedProxClient(AdaptiveDriftConstraintClient):
    anchor_model = "model"
