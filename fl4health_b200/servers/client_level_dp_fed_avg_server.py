"""Server for client-level DP FedAvgM (parity: ``fl4health/servers/client_level_dp_fed_avg_server.py:23-151``): polls
sample counts (needed for weighted aggregation), picks the Poisson / fixed-without-replacement client accountant that
matches the client manager, logs (epsilon, delta)."""

from __future__ import annotations

from collections.abc import Callable, Sequence
from logging import INFO
from math import ceil
from typing import Any

from fl4health_b200.checkpointing.server_module import ClippingBitServerCheckpointAndStateModule
from fl4health_b200.client_managers.fixed_without_replacement_manager import FixedSamplingByFractionClientManager
from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.privacy.fl_accountants import (
    ClientLevelAccountant,
    FlClientLevelAccountantFixedSamplingNoReplacement,
    FlClientLevelAccountantPoissonSampling,
)
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.strategies.client_dp_fedavgm import ClientLevelDPFedAvgM


class ClientLevelDPFedAvgServer(FlServer):
    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        strategy: ClientLevelDPFedAvgM,
        server_noise_multiplier: float,
        num_server_rounds: int,
        checkpoint_and_state_module: ClippingBitServerCheckpointAndStateModule | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        delta: int | None = None,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        transport: Any = None,
    ) -> None:
        if checkpoint_and_state_module is not None:
            assert isinstance(checkpoint_and_state_module, ClippingBitServerCheckpointAndStateModule)
        super().__init__(client_manager=client_manager, fl_config=fl_config, strategy=strategy, reporters=reporters,
                         checkpoint_and_state_module=checkpoint_and_state_module,
                         on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
                         accept_failures=accept_failures, transport=transport)
        self.accountant: ClientLevelAccountant
        self.server_noise_multiplier = server_noise_multiplier
        self.num_server_rounds = num_server_rounds
        self.delta = delta

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        assert isinstance(self.strategy, ClientLevelDPFedAvgM)
        sample_counts = self.poll_clients_for_sample_counts(timeout)
        self.strategy.sample_counts = sample_counts  # weighted aggregation needs the per-client weights
        self.setup_privacy_accountant(sample_counts)
        return super().fit(num_rounds=num_rounds, timeout=timeout)

    def setup_privacy_accountant(self, sample_counts: list[int]) -> None:
        assert isinstance(self.strategy, ClientLevelDPFedAvgM)
        num_clients = len(sample_counts)
        target_delta = 1.0 / num_clients if self.delta is None else self.delta
        if isinstance(self._client_manager, PoissonSamplingClientManager):
            self.accountant = FlClientLevelAccountantPoissonSampling(self.strategy.fraction_fit, self.server_noise_multiplier)
        else:
            assert isinstance(self._client_manager, FixedSamplingByFractionClientManager)
            sampled = ceil(num_clients * self.strategy.fraction_fit)
            self.accountant = FlClientLevelAccountantFixedSamplingNoReplacement(num_clients, sampled, self.server_noise_multiplier)
        epsilon = self.accountant.get_epsilon(self.num_server_rounds, target_delta)
        log(INFO, f"Model privacy after full training will be ({epsilon}, {target_delta})")
