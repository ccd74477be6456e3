"""SCAFFOLD servers (parity: ``fl4health/servers/scaffold_server.py:21-301``): optional warm start — one extra fit
pass over all clients that only updates the control variates (weights are discarded)."""

from __future__ import annotations

from collections.abc import Callable, Sequence
from logging import DEBUG, ERROR, INFO
from typing import Any

from fl4health_b200.checkpointing.server_module import ScaffoldServerCheckpointAndStateModule
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Scalar, ndarrays_to_parameters, parameters_to_ndarrays
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.utils.functions import decode_and_pseudo_sort_results
from fl4health_b200.strategies.scaffold import Scaffold


class ScaffoldServer(FlServer):
    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        strategy: Scaffold,
        reporters: Sequence[BaseReporter] | None = None,
        checkpoint_and_state_module: ScaffoldServerCheckpointAndStateModule | None = None,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        warm_start: bool = False,
        transport: Any = None,
    ) -> None:
        if checkpoint_and_state_module is not None:
            assert isinstance(checkpoint_and_state_module, ScaffoldServerCheckpointAndStateModule), (
                "checkpoint_and_state_module must have type ScaffoldServerCheckpointAndStateModule"
            )
        assert isinstance(strategy, Scaffold)
        super().__init__(
            client_manager=client_manager, fl_config=fl_config, strategy=strategy, reporters=reporters,
            checkpoint_and_state_module=checkpoint_and_state_module,
            on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
            accept_failures=accept_failures, transport=transport,
        )
        self.warm_start = warm_start

    def _get_initial_parameters(self, server_round: int, timeout: float | None):  # noqa: ANN202
        """With ``warm_start`` the initial control variates are estimated by one training pass on every client."""
        assert isinstance(self.strategy, Scaffold)
        initial_parameters = self.strategy.initialize_parameters(client_manager=self._client_manager)
        assert initial_parameters is not None, "Scaffold requires initial parameters (weights ++ variates)"
        if not self.warm_start:
            log(INFO, "Using initial global parameters provided by strategy")
            return initial_parameters
        log(INFO, "Using Warm Start Strategy. Waiting for clients to be available for polling")
        client_instructions = self.strategy.configure_fit_all(
            server_round=0, parameters=initial_parameters, client_manager=self._client_manager
        )
        if not client_instructions:
            log(ERROR, "Warm Start initialization failed: No clients selected")
            return initial_parameters
        log(DEBUG, f"Warm start: strategy sampled {len(client_instructions)} clients")
        results, failures = self.transport.fit_clients(client_instructions, self.max_workers, timeout, group_id=0)
        log(DEBUG, f"Warm Start: Received {len(results)} results and {len(failures)} failures")
        if not results:
            log(ERROR, "Warm Start initialization failed: no client returned a result")
            return initial_parameters
        packer = self.strategy.parameter_packer
        decoded = [arrays for _, arrays, _ in decode_and_pseudo_sort_results(results, materialize=False)]
        _, variate_updates = packer.unpack_parameters(self.strategy.aggregate(decoded))
        # The clients start round 1 from the ORIGINAL weights and the warmed-up variates c0 + (|S|/N) mean(delta c_i).  As
        # in the reference (scaffold_server.py:96-143, where the value is computed but never stored) the strategy's own
        # running variates are NOT advanced by the warm start: round 1's server update starts from the initial ones.
        warmed = self.strategy.compute_updated_parameters(self.strategy.fraction_fit, self.strategy.server_control_variates, variate_updates)
        original_weights, _ = packer.unpack_parameters(parameters_to_ndarrays(initial_parameters))
        return ndarrays_to_parameters(packer.pack_parameters(original_weights, warmed))

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        assert isinstance(self.strategy, Scaffold)
        return super().fit(num_rounds=num_rounds, timeout=timeout)


class DPScaffoldServer(ScaffoldServer):
    """SCAFFOLD with instance-level DP: reports the (epsilon, delta) guarantee before training
    (parity: ``scaffold_server.py:184-301``)."""

    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        noise_multiplier: float,
        batch_size: int,
        num_server_rounds: int,
        strategy: Scaffold,
        local_epochs: int | None = None,
        local_steps: int | None = None,
        delta: float | None = None,
        checkpoint_and_state_module: Any = None,
        warm_start: bool = False,
        reporters: Sequence[BaseReporter] | None = None,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        transport: Any = None,
    ) -> None:
        ScaffoldServer.__init__(
            self, client_manager=client_manager, fl_config=fl_config, strategy=strategy, reporters=reporters,
            checkpoint_and_state_module=checkpoint_and_state_module,
            on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
            accept_failures=accept_failures, warm_start=warm_start, transport=transport,
        )
        from fl4health_b200.servers.instance_level_dp_server import InstanceLevelDpServer

        self._dp = InstanceLevelDpServer.__new__(InstanceLevelDpServer)
        self.noise_multiplier, self.batch_size, self.num_server_rounds = noise_multiplier, batch_size, num_server_rounds
        self.local_epochs, self.local_steps, self.delta = local_epochs, local_steps, delta

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        from fl4health_b200.servers.instance_level_dp_server import InstanceLevelDpServer

        InstanceLevelDpServer.setup_privacy_accountant_and_log(self, timeout)  # type: ignore[arg-type]
        return ScaffoldServer.fit(self, num_rounds=num_rounds, timeout=timeout)
