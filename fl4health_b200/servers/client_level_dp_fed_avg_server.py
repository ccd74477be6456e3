"""Server for client-level DP FedAvgM (parity: ``fl4health/servers/client_level_dp_fed_avg_server.py:23-151``): polls
sample counts (needed for weighted aggregation), picks the Poisson / fixed-without-replacement client accountant that
matches the client manager, logs (epsilon, delta)."""

from __future__ import annotations

from logging import INFO
from math import ceil
from typing import Any

from fl4health_b200.checkpointing.server_module import ClippingBitServerCheckpointAndStateModule
from fl4health_b200.client_managers.fixed_without_replacement_manager import FixedSamplingByFractionClientManager
from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config
from fl4health_b200.privacy.fl_accountants import (
    ClientLevelAccountant,
    FlClientLevelAccountantFixedSamplingNoReplacement,
    FlClientLevelAccountantPoissonSampling,
)
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
        delta: int | None = None,
        **server_options: Any,
    ) -> None:
        """``server_options``: the remaining ``FlServer`` keywords (``reporters``, ``on_init_parameters_config_fn``,
        ``server_name``, ``accept_failures``, ``transport``)."""
        assert checkpoint_and_state_module is None or isinstance(checkpoint_and_state_module, ClippingBitServerCheckpointAndStateModule)
        super().__init__(client_manager, fl_config, strategy, checkpoint_and_state_module=checkpoint_and_state_module,
                         **server_options)
        self.server_noise_multiplier, self.num_server_rounds, self.delta = server_noise_multiplier, num_server_rounds, delta
        self.accountant: ClientLevelAccountant

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        assert isinstance(self.strategy, ClientLevelDPFedAvgM)
        self.strategy.sample_counts = self.poll_clients_for_sample_counts(timeout)  # weights of the noisy aggregate
        self.setup_privacy_accountant(self.strategy.sample_counts)
        return super().fit(num_rounds=num_rounds, timeout=timeout)

    def _accountant_for(self, population: int, sampling_rate: float) -> ClientLevelAccountant:
        """The accountant matching how the client manager samples: Poisson (independent coin flips) or a fixed-size draw
        without replacement."""
        manager = self._client_manager
        if isinstance(manager, PoissonSamplingClientManager):
            return FlClientLevelAccountantPoissonSampling(sampling_rate, self.server_noise_multiplier)
        assert isinstance(manager, FixedSamplingByFractionClientManager)
        return FlClientLevelAccountantFixedSamplingNoReplacement(population, ceil(population * sampling_rate),
                                                                 self.server_noise_multiplier)

    def setup_privacy_accountant(self, sample_counts: list[int]) -> None:
        assert isinstance(self.strategy, ClientLevelDPFedAvgM)
        population = len(sample_counts)
        target_delta = self.delta if self.delta is not None else 1.0 / population
        self.accountant = self._accountant_for(population, self.strategy.fraction_fit)
        epsilon = self.accountant.get_epsilon(self.num_server_rounds, target_delta)
        log(INFO, f"Model privacy after full training will be ({epsilon}, {target_delta})")
