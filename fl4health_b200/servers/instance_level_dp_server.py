"""Server for instance-level DP FL (parity: ``fl4health/servers/instance_level_dp_server.py:19-168``): polls clients for
sample counts, builds the ``FlInstanceLevelAccountant`` and logs (epsilon, delta) before training starts."""

from __future__ import annotations

from collections.abc import Callable, Sequence
from logging import INFO
from math import ceil
from typing import Any

from fl4health_b200.checkpointing.server_module import OpacusServerCheckpointAndStateModule
from fl4health_b200.client_managers.poisson_sampling_manager import PoissonSamplingClientManager
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar
from fl4health_b200.privacy.fl_accountants import FlInstanceLevelAccountant
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg


class InstanceLevelDpServer(FlServer):
    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        noise_multiplier: float,
        batch_size: int,
        num_server_rounds: int,
        strategy: BasicFedAvg,
        local_epochs: int | None = None,
        local_steps: int | None = None,
        checkpoint_and_state_module: OpacusServerCheckpointAndStateModule | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        delta: float | None = None,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        transport: Any = None,
    ) -> None:
        if checkpoint_and_state_module is not None:
            assert isinstance(checkpoint_and_state_module, OpacusServerCheckpointAndStateModule)
        super().__init__(client_manager=client_manager, fl_config=fl_config, strategy=strategy, reporters=reporters,
                         checkpoint_and_state_module=checkpoint_and_state_module,
                         on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
                         accept_failures=accept_failures, transport=transport)
        assert (local_epochs is None) != (local_steps is None), "Either local_epochs or local_steps should be set but not both"
        self.accountant: FlInstanceLevelAccountant
        self.local_epochs, self.local_steps = local_epochs, local_steps
        self.noise_multiplier, self.batch_size, self.num_server_rounds, self.delta = noise_multiplier, batch_size, num_server_rounds, delta

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        self.setup_privacy_accountant_and_log(timeout)
        return super().fit(num_rounds=num_rounds, timeout=timeout)

    def setup_privacy_accountant(self, sample_counts: list[int]) -> None:
        """Build the FL accountant from the polled per-client training-set sizes (parity:
        ``instance_level_dp_server.py:131-169``).  With a step budget the epoch count charged is the largest any client
        can reach (ceil(steps * batch / smallest dataset)), so the privacy loss is never under-estimated."""
        assert isinstance(self._client_manager, PoissonSamplingClientManager), "instance-level DP requires Poisson client sampling"
        if self.local_epochs is not None:
            epochs_per_round = self.local_epochs
        else:
            assert self.local_steps is not None
            epochs_per_round = max(ceil(self.local_steps * self.batch_size / max(min(sample_counts), 1)), 1)
        self.accountant = FlInstanceLevelAccountant(
            client_sampling_rate=self.strategy.fraction_fit, noise_multiplier=self.noise_multiplier,
            epochs_per_round=epochs_per_round, client_batch_sizes=[self.batch_size] * len(sample_counts),
            client_dataset_sizes=sample_counts,
        )

    def setup_privacy_accountant_and_log(self, timeout: float | None) -> None:
        """Also used (unbound) by ``DPScaffoldServer``."""
        sample_counts = self.poll_clients_for_sample_counts(timeout)
        total_samples = sum(sample_counts)
        InstanceLevelDpServer.setup_privacy_accountant(self, sample_counts)
        target_delta = self.delta if self.delta is not None else 1.0 / total_samples
        epsilon = self.accountant.get_epsilon(self.num_server_rounds, target_delta)
        log(INFO, f"Model privacy after full training will be ({epsilon}, {target_delta})")
