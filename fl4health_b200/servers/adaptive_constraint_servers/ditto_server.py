"""Ditto server (parity: ``fl4health/servers/adaptive_constraint_servers/ditto_server.py``): an ``FlServer`` that insists on the
``FedAvgWithAdaptiveConstraint`` strategy and, when checkpointing, on the module that strips the packed mu."""

from __future__ import annotations

from collections.abc import Callable, Sequence
from typing import Any

from fl4health_b200.checkpointing.server_module import AdaptiveConstraintServerCheckpointAndStateModule
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.strategies.fedavg_with_adaptive_constraint import FedAvgWithAdaptiveConstraint


class DittoServer(FlServer):
    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        strategy: FedAvgWithAdaptiveConstraint,
        reporters: Sequence[BaseReporter] | None = None,
        checkpoint_and_state_module: AdaptiveConstraintServerCheckpointAndStateModule | None = None,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        transport: Any = None,
    ) -> None:
        assert isinstance(strategy, FedAvgWithAdaptiveConstraint), (
            "Strategy must be of base type FedAvgWithAdaptiveConstraint"
        )
        if checkpoint_and_state_module is not None:
            assert isinstance(checkpoint_and_state_module, AdaptiveConstraintServerCheckpointAndStateModule), (
                "checkpoint_and_state_module must have type AdaptiveConstraintServerCheckpointAndStateModule"
            )
        super().__init__(
            client_manager=client_manager, fl_config=fl_config, strategy=strategy, reporters=reporters,
            checkpoint_and_state_module=checkpoint_and_state_module,
            on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
            accept_failures=accept_failures, transport=transport,
        )
