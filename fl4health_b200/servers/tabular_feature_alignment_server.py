"""Server that aligns tabular feature spaces across clients before training (parity:
``fl4health/servers/tabular_feature_alignment_server.py:29-200``).

Before round 1 it runs up to two ``get_properties`` polls: (1) if it has no schema ("source of truth") it asks ONE
randomly chosen client for its schema; (2) it broadcasts the schema through the fit config, lets every client align,
and asks one client for the aligned input / output dimensions, from which ``initialize_parameters(in, out)`` builds the
global model's initial parameters.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
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


class _RoundConfig:
    """What the strategy hands to clients each round: the server's FL config (which, once the schema is known, carries
    it under ``FEATURE_INFO``) stamped with the round and with whether a schema has been agreed yet.  One object for
    fit and evaluate; the handshake flips ``schema_agreed`` instead of re-installing closures."""

    def __init__(self, fl_config: Config) -> None:
        self.fl_config, self.schema_agreed = fl_config, False

    def __call__(self, current_server_round: int) -> Config:
        return fit_config(self.fl_config, self.schema_agreed, current_server_round)


class TabularFeatureAlignmentServer(FlServer):
    """Pre-training handshake, in two polls of ``get_properties`` (both skipped on later ``fit`` calls):

    1. *schema*: taken from ``tabular_features_source_of_truth`` or, failing that, from ONE client drawn with the
       server-side sampling stream; published to all clients through the round config;
    2. *dimensions*: every client aligns its frame to the schema and reports the aligned input / output widths;
       ``initialize_parameters(in, out)`` then builds the global model the first round starts from."""

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
        for overwritten in ("on_fit_config_fn", "initial_parameters"):
            if getattr(strategy, overwritten) is not None:
                log(WARNING, f"strategy.{overwritten} will be overwritten.")
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
        self.dimension_info: dict[str, int] = {}
        self._round_config = _RoundConfig(self.fl_config)
        self.strategy.on_fit_config_fn = self._round_config
        if getattr(self.strategy, "on_evaluate_config_fn", None) is None:  # a user-supplied evaluation config is kept
            self.strategy.on_evaluate_config_fn = self._round_config

    @property
    def source_info_gathered(self) -> bool:
        return self._round_config.schema_agreed

    def _set_dimension_info(self, input_dimension: int, output_dimension: int) -> None:
        self.dimension_info.update({INPUT_DIMENSION: input_dimension, OUTPUT_DIMENSION: output_dimension})

    def _get_initial_parameters(self, server_round: int, timeout: float | None) -> Parameters:
        assert {INPUT_DIMENSION, OUTPUT_DIMENSION} <= set(self.dimension_info), "the dimension poll has not run"
        return self.initialize_parameters(self.dimension_info[INPUT_DIMENSION], self.dimension_info[OUTPUT_DIMENSION])

    def _negotiate(self, timeout: float | None) -> None:
        if self.tab_features_info is not None:
            log(INFO, "Features information source already specified. Sending to clients to perform feature alignment.")
            schema = self.tab_features_info.to_json()
        else:
            schema = self.poll_clients_for_feature_info(timeout)
        self.fl_config[FEATURE_INFO] = schema
        self._round_config.schema_agreed = True
        dimensions = self.poll_clients_for_dimension_info(timeout)
        log(DEBUG, f"input dimension: {dimensions[0]}, output dimension: {dimensions[1]}")
        self._set_dimension_info(*dimensions)
        self.initial_polls_complete = True

    def fit(self, num_rounds: int, timeout: float | None = None) -> tuple[History, float]:
        if not self.initial_polls_complete:
            self._negotiate(timeout)
        return super().fit(num_rounds=num_rounds, timeout=timeout)

    def _poll(self, who: str, timeout: float | None) -> list:
        instructions = self.strategy.configure_poll(server_round=1, client_manager=self._client_manager)
        if who == "one":  # one client's schema becomes the source of truth
            instructions = sampling_streams.python.sample(population=instructions, k=1)
        answers, _ = self.transport.poll_clients(instructions, max_workers=self.max_workers, timeout=timeout)
        return [response.properties for _, response in answers]

    def poll_clients_for_feature_info(self, timeout: float | None) -> str:
        log(INFO, "Feature information source unspecified. Polling clients for feature information.")
        (properties,) = self._poll("one", timeout)
        return str(properties[FEATURE_INFO])

    def poll_clients_for_dimension_info(self, timeout: float | None) -> tuple[int, int]:
        log(INFO, "Waiting for Clients to align features and then polling for dimension information.")
        answers = self._poll("all", timeout)  # every client aligns (all are set up before round 1); the widths are identical
        assert answers, "no client answered the dimension poll"
        return int(answers[0][INPUT_DIMENSION]), int(answers[0][OUTPUT_DIMENSION])
