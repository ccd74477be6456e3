"""Servers shared by the research experiments.

* ``PersonalServer`` — for methods whose only meaningful model lives on the clients (APFL, FENDA, Ditto, MR-MTL …):
  no global checkpoint, but the weighted aggregated validation loss is tracked for hyper-parameter selection
  (parity: ``research/cifar10/personal_server.py:22-65``).
* ``FullExchangeServer`` — for methods that exchange the complete model (FedAvg, FedProx, FedAdam, SCAFFOLD …): the
  global model is checkpointed (best aggregated loss and latest) by hydrating it through a full exchanger (parity:
  ``research/flamby/flamby_servers/full_exchange_server.py``).
"""

from __future__ import annotations

from logging import INFO
from pathlib import Path
from typing import Any

from torch import nn

from fl4health_b200.checkpointing.checkpointer import BestLossTorchModuleCheckpointer, LatestTorchModuleCheckpointer
from fl4health_b200.checkpointing.server_module import BaseServerCheckpointAndStateModule
from fl4health_b200.common.logger import log
from fl4health_b200.parameter_exchange.full_exchanger import FullParameterExchanger
from fl4health_b200.servers.base_server import FlServer


class _BestLossTracking:
    """Mixin: remember the best weighted validation loss seen in any round (what the HP sweep ranks by)."""

    best_aggregated_loss: float | None = None

    def _track(self, loss_aggregated: float | None) -> None:
        if loss_aggregated is None:
            return
        if self.best_aggregated_loss is None or loss_aggregated <= self.best_aggregated_loss:
            log(INFO, f"Best aggregated loss so far: {loss_aggregated} (previous: {self.best_aggregated_loss})")
            self.best_aggregated_loss = float(loss_aggregated)


class PersonalServer(_BestLossTracking, FlServer):
    def __init__(self, client_manager: Any, fl_config: dict[str, Any], strategy: Any = None, **kwargs: Any) -> None:
        kwargs.pop("checkpoint_and_state_module", None)  # there is no global model to checkpoint
        super().__init__(client_manager, fl_config, strategy, checkpoint_and_state_module=None, **kwargs)

    def evaluate_round(self, server_round: int, timeout: float | None) -> Any:
        outcome = super().evaluate_round(server_round, timeout)
        assert outcome is not None, "personal methods rank hyper-parameters by the federated validation loss"
        self._track(outcome[0])
        return outcome


def make_personal(server_cls: type) -> type:
    """``PersonalServer`` behaviour on top of a method-specific server class (Ditto / MR-MTL / FedProx servers)."""
    if issubclass(server_cls, _BestLossTracking):
        return server_cls

    def evaluate_round(self: Any, server_round: int, timeout: float | None) -> Any:
        outcome = server_cls.evaluate_round(self, server_round, timeout)
        if outcome is not None:
            self._track(outcome[0])
        return outcome

    return type(f"Personal{server_cls.__name__}", (_BestLossTracking, server_cls), {"evaluate_round": evaluate_round})


class FullExchangeServer(_BestLossTracking, FlServer):
    def __init__(self, client_manager: Any, fl_config: dict[str, Any], strategy: Any = None, model: nn.Module | None = None,
                 checkpoint_dir: Path | None = None, **kwargs: Any) -> None:
        module = None
        if model is not None and checkpoint_dir is not None:
            checkpoint_dir.mkdir(parents=True, exist_ok=True)
            module = BaseServerCheckpointAndStateModule(
                model=model, parameter_exchanger=FullParameterExchanger(),
                model_checkpointers=[BestLossTorchModuleCheckpointer(str(checkpoint_dir), "server_best_model.pkl"),
                                     LatestTorchModuleCheckpointer(str(checkpoint_dir), "server_last_model.pkl")])
        super().__init__(client_manager, fl_config, strategy, checkpoint_and_state_module=module, **kwargs)

    def evaluate_round(self, server_round: int, timeout: float | None) -> Any:
        outcome = super().evaluate_round(server_round, timeout)
        if outcome is not None:
            self._track(outcome[0])
        return outcome
