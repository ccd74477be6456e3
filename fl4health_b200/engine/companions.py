"""Companion models: extra ``nn.Module``s a client keeps next to ``self.model``.

Many FL algorithms are "the basic client plus one more copy of the network": Ditto trains a *global* twin that is the
one exchanged with the server, MR-MTL / FedProx-style methods keep a frozen *anchor* holding the last aggregate, MOON
and PerFCL keep frozen snapshots for their contrastive terms.  The reference re-implements the plumbing for each
(``setup_client`` override to build and place the copy, ``update_before_train`` / ``validate`` overrides to flip its
``train()`` / ``eval()`` state, bespoke optimizer look-ups; e.g. ``fl4health/clients/ditto_client.py:150-215``,
``mr_mtl_client.py:80-140``).  Here a client *declares* its companions:

    companions = {"global_model": Companion(factory="get_global_model", trainable=True, mode=FOLLOW)}

and ``BasicClient`` does the rest: builds each companion with the named factory method at set-up, places it like the
main model (device, flat arena, channels-last, bf16 shadow when trainable), offers its arena to the optimizer
translation, and keeps its training mode in step with the phase of the round.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from torch import nn

from fl4health_b200.engine.modes import set_training

FOLLOW = "follow"   # train() while the client trains, eval() while it evaluates (a second trained network)
FROZEN = "frozen"   # always eval(), parameters do not require gradients (anchors, snapshots)


@dataclass(frozen=True)
class Companion:
    factory: str            # name of the client method ``(config) -> nn.Module`` that builds it
    trainable: bool = True  # gets a gradient region / may own an optimizer
    mode: str = FOLLOW


def build_companions(client: Any, config: dict) -> None:
    """Instantiate and place every declared companion as an attribute of ``client`` (before the main model, so user
    factories may share construction code with ``get_model``)."""
    for name, spec in getattr(client, "companions", {}).items():
        module = client._place_model(getattr(client, spec.factory)(config), with_grad=spec.trainable)
        if spec.mode == FROZEN:
            for param in module.parameters():
                param.requires_grad = False
        setattr(client, name, module)


def companion_modules(client: Any, trainable_only: bool = False) -> list[nn.Module]:
    specs = getattr(client, "companions", {})
    return [getattr(client, name) for name, spec in specs.items()
            if hasattr(client, name) and (spec.trainable or not trainable_only)]


def set_phase(client: Any, training: bool) -> None:
    """Put the main model and every companion in the mode the current phase (local training / evaluation) calls for."""
    set_training(client.model, training)
    for name, spec in getattr(client, "companions", {}).items():
        module = getattr(client, name, None)
        if module is not None:
            set_training(module, training and spec.mode == FOLLOW)
