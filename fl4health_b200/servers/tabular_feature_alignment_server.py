"""Server that aligns tabular feature spaces across clients before training (parity:
``fl4health/servers/tabular_feature_alignment_server.py:29-200``).

Before round 1 it runs up to two ``get_properties`` polls: (1) if it has no schema ("source of truth") it asks ONE
randomly chosen client for its schema; (2) it broadcasts the schema through the fit config, lets every client align,
and asks one client for the aligned input / output dimensions, from which ``initialize_parameters(in, out)`` builds the
global model's initial parameters.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from functools import partial
from logging import DEBUG, INFO, WARNING

from fl4health_b200.checkpointing.server_module import BaseServerCheckpointAndStateModule
from fl4health_b200.common.history import History
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Config, Parameters, Scalar
from fl4health_b200.feature_alignment.constants import (
    CURRENT_SERVER_ROUND,
    FEATURE_INFO,
    INPUT_DIMENSION,
    OUTPUT_DIMENSION,
    SOURCE_SPECIFIED,
)
from fl4health_b200.feature_alignment.tab_features_info_encoder import TabularFeaturesInfoEncoder
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import ClientManager, sampling_streams
from fl4health_b200.strategies.basic_fedavg import BasicFedAvg


def fit_config(config: Config, source_specified: bool, current_server_round: int) -> Config:
    config[SOURCE_SPECIFIED] = source_specified
    config[CURRENT_SERVER_ROUND] = current_server_round
    return config


class TabularFeatureAlignmentServer(FlServer):
    def __init__(
        self,
        client_manager: ClientManager,
        config: Config,
        initialize_parameters: Callable[[int, int], Parameters],
        strategy: BasicFedAvg,
        tabular_features_source_of_truth: TabularFeaturesInfoEncoder | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        checkpoint_and_state_module: BaseServerCheckpointAndStateModule | None = None,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]] | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
    ) -> None:
        if strategy.on_fit_config_fn is not None:
            log(WARNING, "strategy.on_fit_config_fn will be overwritten.")
        if strategy.initial_parameters is not None:
            log(WARNING, "strategy.initial_parameters will be overwritten.")
        super().__init__(
            client_manager=client_manager, fl_config=config, strategy=strategy, reporters=reporters,
            checkpoint_and_state_module=checkpoint_and_state_module,
            on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
            accept_failures=accept_failures,
        )
        assert isinstance(self.strategy, BasicFedAvg), "This server is only compatible with BasicFedAvg at this time"
        self.initial_polls_complete = False
        self.tab_features_info = tabular_features_source_of_truth
        self.initialize_parameters = initialize_parameters
        self.source_info_gathered = False
        self.dimension_info: dict[str, int] = {}
        self._install_config_fn()

    def _install_config_fn(self) -> None:
        fn = partial(fit_config, self.fl_config, self.source_info_gathered)
        self.strategy.on_fit_config_fn = fn
        if getattr(self.strategy, "on_evaluate_config_fn", None) is None or getattr(self, "_owns_eval_config", False):
            self.strategy.on_evaluate_config_fn = fn
            self._owns_eval_config = True

    def _set_dimension_info(self, input_dimension: int, output_dimension: int) -> None:
        self.dimension_info[INPUT_DIMENSION] = input_dimension
        self.dimension_info[OUTPUT_DIMENSION] = output_dimension

    def _get_initial_parameters(self, server_round: int, timeout: float | None) -> Parameters:
        assert INPUT_DIMENSION in self.dimension_info and OUTPUT_DIMENSION in self.dimension_info
        return self.initialize_parameters(self.dimension_info[INPUT_DIMENSION], self.dimension_info[OUTPUT_DIMENSION])

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        if not self.initial_polls_complete:
            if self.tab_features_info is None:
                feature_info_source = self.poll_clients_for_feature_info(timeout)
            else:
                log(INFO, "Features information source already specified. Sending to clients to perform feature alignment.")
                feature_info_source = self.tab_features_info.to_json()
            self.fl_config[FEATURE_INFO] = feature_info_source
            self.source_info_gathered = True
            self._install_config_fn()
            input_dimension, output_dimension = self.poll_clients_for_dimension_info(timeout)
            log(DEBUG, f"input dimension: {input_dimension}, output dimension: {output_dimension}")
            self._set_dimension_info(input_dimension, output_dimension)
            self.initial_polls_complete = True
        return super().fit(num_rounds=num_rounds, timeout=timeout)

    def poll_clients_for_feature_info(self, timeout: float | None) -> str:
        log(INFO, "Feature information source unspecified. Polling clients for feature information.")
        instructions = self.strategy.configure_poll(server_round=1, client_manager=self._client_manager)
        chosen = sampling_streams.python.sample(population=instructions, k=1)  # one client's schema becomes the source of truth
        results, _ = self.transport.poll_clients(chosen, max_workers=self.max_workers, timeout=timeout)
        assert len(results) == 1
        return str(results[0][1].properties[FEATURE_INFO])

    def poll_clients_for_dimension_info(self, timeout: float | None) -> tuple[int, int]:
        log(INFO, "Waiting for Clients to align features and then polling for dimension information.")
        instructions = self.strategy.configure_poll(server_round=1, client_manager=self._client_manager)
        # every client aligns (so all are set up before round 1); dimensions are identical, read the first answer
        results, _ = self.transport.poll_clients(instructions, max_workers=self.max_workers, timeout=timeout)
        assert len(results) >= 1
        properties = results[0][1].properties
        return int(properties[INPUT_DIMENSION]), int(properties[OUTPUT_DIMENSION])
