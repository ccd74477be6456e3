"""Frozen reference networks for contrastive / distillation terms.

MOON, PerFCL (and anything distilling from "the model as it was at some earlier point") keep frozen copies of (parts
of) the trained network and run them under ``no_grad`` next to the live forward.  ``SnapshotBank`` owns those copies:

* ``capture(slot, module)``   freeze a deep copy into ``slot``; a slot created with ``keep=n`` is a FIFO of the last n;
* ``get(slot)`` / ``all(slot)``  the newest copy / every copy of the slot, oldest first;
* ``variant()``               a hashable tag of the current contents — what the CUDA-graph step runner needs to know,
                              because a captured step is only valid for the exact frozen tensors it was recorded with.

Reference: ad-hoc attributes + ``clone_and_freeze_model`` calls in ``fl4health/clients/moon_client.py:150-200`` and
``perfcl_client.py:190-240``.
"""

from __future__ import annotations

from collections import deque
from collections.abc import Hashable

from torch import nn

from fl4health_b200.utils.client import clone_and_freeze_model


class SnapshotBank:
    def __init__(self, **slots: int) -> None:
        """``SnapshotBank(old=3, anchor=1)``: slot name -> how many generations to keep."""
        self._slots: dict[str, deque[nn.Module]] = {name: deque(maxlen=keep) for name, keep in slots.items()}

    def capture(self, slot: str, module: nn.Module) -> nn.Module:
        frozen = clone_and_freeze_model(module)
        self._slots[slot].append(frozen)
        return frozen

    def get(self, slot: str) -> nn.Module | None:
        held = self._slots[slot]
        return held[-1] if held else None

    def all(self, slot: str) -> list[nn.Module]:
        return list(self._slots[slot])

    def filled(self, *slots: str) -> bool:
        return all(len(self._slots[name]) > 0 for name in (slots or self._slots))

    def resize(self, slot: str, keep: int) -> None:
        self._slots[slot] = deque(self._slots[slot], maxlen=keep)

    def variant(self) -> Hashable:
        return tuple((name, tuple(id(module) for module in held)) for name, held in self._slots.items())
