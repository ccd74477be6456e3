"""nnU-Net federation server (parity: ``fl4health/servers/nnunet_server.py:31-264``).

Before round 1 (``update_before_fit``) one random client is polled through ``get_properties`` when (a) no global
``nnunet_plans`` were put in the config — that client plans the experiment on its local dataset and its plans become
the federation's — and/or (b) server-side checkpointing needs the architecture (input channels / segmentation heads
come from a client's dataset).  The plans are then injected into every fit / evaluate config.

The network constructor is pluggable (``model_builder``); by default nnunetv2's trainer builds it (imported lazily:
the optional dependency is only needed when a server-side model is requested).
"""

from __future__ import annotations

import pickle
from collections.abc import Callable, Sequence
from logging import INFO
from typing import Any

from torch import nn

from fl4health_b200.checkpointing.server_module import NnUnetServerCheckpointAndStateModule
from fl4health_b200.common.logger import log
from fl4health_b200.common.typing import Code, Config, EvaluateIns, FitIns, GetPropertiesIns, Parameters, Scalar
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.servers.base_server import FlServer
from fl4health_b200.servers.client_manager import ClientManager
from fl4health_b200.servers.client_proxy import ClientProxy
from fl4health_b200.strategies.strategy import Strategy
from fl4health_b200.utils.config import narrow_dict_type
from fl4health_b200.utils.nnunet_utils import NnunetConfig

FIT_CFG_FN = Callable[[int, Parameters, ClientManager], list[tuple[ClientProxy, FitIns]]]
EVAL_CFG_FN = Callable[[int, Parameters, ClientManager], list[tuple[ClientProxy, EvaluateIns]]]
CFG_FN = FIT_CFG_FN | EVAL_CFG_FN

CFG_FN = Callable[..., Any]
ModelBuilder = Callable[[dict, NnunetConfig, int, int, bool], nn.Module]


def add_items_to_config_fn(fn: CFG_FN, items: Config) -> CFG_FN:
    """Wrap ``strategy.configure_fit/evaluate`` so every instruction's config also carries ``items``."""

    def new_fn(*args: Any, **kwargs: Any) -> Any:
        instructions = fn(*args, **kwargs)
        for _, ins in instructions:
            ins.config.update(items)
        return instructions

    return new_fn


def nnunetv2_model_builder(plans: dict, config: NnunetConfig, num_input_channels: int, num_segmentation_heads: int,
                           deep_supervision: bool) -> nn.Module:
    try:
        from nnunetv2.training.nnUNetTrainer.nnUNetTrainer import nnUNetTrainer  # type: ignore[import-not-found]
        from nnunetv2.utilities.plans_handling.plans_handler import PlansManager  # type: ignore[import-not-found]
    except ImportError as exc:
        raise ImportError("Building the server-side nnU-Net needs the optional 'nnunetv2' package (or pass model_builder=...)") from exc
    configuration = PlansManager(plans).get_configuration(config.value)
    return nnUNetTrainer.build_network_architecture(
        configuration.network_arch_class_name, configuration.network_arch_init_kwargs,
        configuration.network_arch_init_kwargs_req_import, num_input_channels, num_segmentation_heads, deep_supervision,
    )


def _trainer_model_builder(trainer_class: type) -> ModelBuilder:
    def build(plans: dict, config: NnunetConfig, num_input_channels: int, num_segmentation_heads: int, deep_supervision: bool) -> nn.Module:
        from nnunetv2.utilities.plans_handling.plans_handler import PlansManager  # type: ignore[import-not-found]

        configuration = PlansManager(plans).get_configuration(config.value)
        return trainer_class.build_network_architecture(
            configuration.network_arch_class_name, configuration.network_arch_init_kwargs,
            configuration.network_arch_init_kwargs_req_import, num_input_channels, num_segmentation_heads, deep_supervision,
        )

    return build


class NnunetServer(FlServer):
    def __init__(
        self,
        client_manager: ClientManager,
        fl_config: Config,
        on_init_parameters_config_fn: Callable[[int], dict[str, Scalar]],
        strategy: Strategy | None = None,
        reporters: Sequence[BaseReporter] | None = None,
        checkpoint_and_state_module: NnUnetServerCheckpointAndStateModule | None = None,
        server_name: str | None = None,
        accept_failures: bool = True,
        model_builder: ModelBuilder = nnunetv2_model_builder,
        global_deep_supervision: bool = False,
        nnunet_trainer_class: type | None = None,
    ) -> None:
        """``nnunet_trainer_class``: an ``nnUNetTrainer`` subclass whose ``build_network_architecture`` builds the global
        model (the reference's knob, ``nnunet_server.py:54-110``); ignored when a custom ``model_builder`` is given."""
        if checkpoint_and_state_module is not None:
            assert isinstance(checkpoint_and_state_module, NnUnetServerCheckpointAndStateModule), (
                "checkpoint_and_state_module must have type NnUnetServerCheckpointAndStateModule")
        else:
            checkpoint_and_state_module = NnUnetServerCheckpointAndStateModule()
        super().__init__(
            client_manager=client_manager, fl_config=fl_config, strategy=strategy, reporters=reporters,
            checkpoint_and_state_module=checkpoint_and_state_module,
            on_init_parameters_config_fn=on_init_parameters_config_fn, server_name=server_name,
            accept_failures=accept_failures,
        )
        if nnunet_trainer_class is not None and model_builder is nnunetv2_model_builder:
            model_builder = _trainer_model_builder(nnunet_trainer_class)
        self.model_builder = model_builder
        self.global_deep_supervision = global_deep_supervision
        self.nnunet_config = NnunetConfig(self.fl_config["nnunet_config"])
        self.nnunet_plans_bytes: bytes
        self.num_input_channels: int
        self.num_segmentation_heads: int

    def initialize_server_model(self) -> None:
        plans = pickle.loads(self.nnunet_plans_bytes)
        self.checkpoint_and_state_module.model = self.model_builder(
            plans, self.nnunet_config, self.num_input_channels, self.num_segmentation_heads, self.global_deep_supervision)

    def update_before_fit(self, num_rounds: int, timeout: float | None) -> None:
        config = self.strategy.configure_fit(0, Parameters([], "None"), self._client_manager)[0][1].config
        plans_bytes = config.get("nnunet_plans")
        module = self.checkpoint_and_state_module
        checkpointer_exists = module.state_checkpointer is not None or module.model_checkpointers is not None
        if not checkpointer_exists and plans_bytes is not None:
            return
        log(INFO, "[PRE-INIT] Requesting properties from one random client via get_properties")
        if plans_bytes is None:
            log(INFO, "\tThis client will be asked to initialize the global nnunet plans")
        if checkpointer_exists:
            log(INFO, "\tThis client's local dataset will be used to determine the number of input and output channels")
        client = self._client_manager.sample(1)[0]
        # through the transport: under SPMD the rank that hosts this client answers and every rank receives the reply
        answered, failed = self.transport.poll_clients([(client, GetPropertiesIns(config=config))], None, timeout)
        if failed or not answered or answered[0][1].status.code != Code.OK:
            raise RuntimeError("Failed to successfully receive properties from client")
        response = answered[0][1]
        properties = response.properties
        self.nnunet_plans_bytes = narrow_dict_type(properties, "nnunet_plans", bytes) if plans_bytes is None else plans_bytes
        assert isinstance(self.nnunet_plans_bytes, bytes)
        self.num_segmentation_heads = narrow_dict_type(properties, "num_segmentation_heads", int)
        self.num_input_channels = narrow_dict_type(properties, "num_input_channels", int)
        if checkpointer_exists:
            self.initialize_server_model()
        extra = {"nnunet_plans": self.nnunet_plans_bytes}
        self.strategy.configure_fit = add_items_to_config_fn(self.strategy.configure_fit, extra)  # type: ignore[method-assign]
        self.strategy.configure_evaluate = add_items_to_config_fn(self.strategy.configure_evaluate, extra)  # type: ignore[method-assign]
        init_fn = self.on_init_parameters_config_fn
        if init_fn is not None:  # the client asked for initial parameters must be able to set itself up too
            self.on_init_parameters_config_fn = lambda server_round: {**init_fn(server_round), **extra}

    def _save_server_state(self) -> None:
        assert getattr(self, "nnunet_plans_bytes", None) is not None and getattr(self, "num_input_channels", None) is not None
        super()._save_server_state()
