"""FedPM server: resets the Beta priors every ``reset_frequency`` rounds (parity: ``fl4health/servers/fedpm_server.py:14-89``)."""

from __future__ import annotations

from collections.abc import Callable, Sequence
from logging import INFO
from typing import Any

from fl4health_b200.checkpointing.server_module import LayerNamesServerCheckpointAndStateModule
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.strategies.fedpm import FedPm


class FedPmServer(FlServer):
    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        strategy: FedPm,
        reporters: Sequence[BaseReporter] | None = None,
        checkpoint_and_state_module: LayerNamesServerCheckpointAndStateModule | None = None,
        reset_frequency: int = 1,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        transport: Any = None,
    ) -> None:
        if checkpoint_and_state_module is not None:
            assert isinstance(checkpoint_and_state_module, LayerNamesServerCheckpointAndStateModule)
        super().__init__(client_manager=client_manager, fl_config=fl_config, strategy=strategy, reporters=reporters,
                         checkpoint_and_state_module=checkpoint_and_state_module,
                         on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
                         accept_failures=accept_failures, transport=transport)
        self.reset_frequency = reset_frequency

    def fit_round(self, server_round: int, timeout: float | None):  # noqa: ANN201
        assert isinstance(self.strategy, FedPm)
        if self.strategy.bayesian_aggregation and server_round > 1 and server_round % self.reset_frequency == 0:
            log(INFO, f"Resetting the Beta priors at the start of round {server_round}")
            self.strategy.reset_beta_priors()
        return super().fit_round(server_round, timeout)
