"""MR-MTL (mean-regularised multi-task learning): only a personal model is trained; the server aggregate is kept in
``initial_global_model`` purely as the drift anchor (parity: ``fl4health/clients/mr_mtl_client.py:18-167``).

The whole variant is a declaration on top of ``AdaptiveDriftConstraintClient``: a frozen companion network receives every
aggregate and is the anchor; the personal ``self.model`` is never overwritten (not even in round 1) and is what is sent
back."""

from __future__ import annotations

from torch import nn

from fl4health_b200.clients.adaptive_drift_constraint_client import AdaptiveDriftConstraintClient
from fl4health_b200.engine.companions import FROZEN, Companion


class MrMtlClient(AdaptiveDriftConstraintClient):
    companions = {"initial_global_model": Companion(factory="get_model", trainable=False, mode=FROZEN)}
    receives_into = "initial_global_model"
    anchor_model = "initial_global_model"

    initial_global_model: nn.Module
