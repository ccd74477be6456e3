"""Loss containers and meters.

Parity: ``fl4health/utils/losses.py:10-234`` (``TrainingLosses``, ``EvaluationLosses``, ``LossMeter``), with a
B200-first change of mechanics: the reference keeps a Python list of per-step loss objects and aggregates through
``torch.FloatTensor([...])`` (a device→host sync per element, SURVEY hot-op L16).  Here a meter owns *device-side
running sums*; ``update`` is a handful of in-place adds that are CUDA-graph capturable and the host only reads the
result once, in ``compute().as_dict()``.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Generic, TypeVar

import torch


def read_scalars(tensors: dict[str, torch.Tensor]) -> dict[str, float]:
    """Host values of several 0-d tensors with ONE device→host transfer (a ``.item()`` per entry costs a copy and a
    stream synchronisation each; right after a local-training phase the GPU sits idle for all of them)."""
    values = list(tensors.values())
    if len(values) > 1 and all(v.is_cuda and v.device == values[0].device for v in values):
        host = torch.stack([v.detach().reshape(()).to(torch.float64) for v in values]).tolist()
        return dict(zip(tensors.keys(), host))
    return {key: float(val.item()) for key, val in tensors.items()}


class Losses(ABC):
    def __init__(self, additional_losses: dict[str, torch.Tensor] | None = None) -> None:
        self.additional_losses: dict[str, torch.Tensor] = additional_losses if additional_losses else {}

    def as_dict(self) -> dict[str, float]:
        return read_scalars(self._named_scalars())

    def _named_scalars(self) -> dict[str, torch.Tensor]:
        """The tensors ``as_dict`` reports, under their reporting keys."""
        return dict(self.additional_losses)

    def _flat_items(self) -> dict[str, torch.Tensor]:
        """All scalar tensors of this container under unique keys (used by the meter)."""
        return {f"additional::{k}": v for k, v in self.additional_losses.items()}

    def detach(self) -> Losses:
        """Same values without autograd history.  The engine hands *detached* losses back to the training loop so no
        autograd graph (and none of its AccumulateGrad nodes) outlives the step — a requirement for CUDA-graph
        capture of the next step and a memory saving otherwise."""
        import copy

        clone = copy.copy(self)
        for name, value in list(vars(clone).items()):
            if isinstance(value, torch.Tensor):
                setattr(clone, name, value.detach())
            elif isinstance(value, dict):
                setattr(clone, name, {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in value.items()})
        return clone

    @staticmethod
    @abstractmethod
    def aggregate(loss_meter: LossMeter) -> Losses:
        raise NotImplementedError


class EvaluationLosses(Losses):
    def __init__(self, checkpoint: torch.Tensor, additional_losses: dict[str, torch.Tensor] | None = None) -> None:
        super().__init__(additional_losses)
        self.checkpoint = checkpoint

    def _named_scalars(self) -> dict[str, torch.Tensor]:
        out = super()._named_scalars()
        out["checkpoint"] = self.checkpoint
        return out

    def _flat_items(self) -> dict[str, torch.Tensor]:
        items = super()._flat_items()
        items["checkpoint"] = self.checkpoint
        return items

    @staticmethod
    def aggregate(loss_meter: LossMeter[EvaluationLosses]) -> EvaluationLosses:
        sums = loss_meter._reduced()
        additional = {k.split("::", 1)[1]: v for k, v in sums.items() if k.startswith("additional::")}
        return EvaluationLosses(checkpoint=sums["checkpoint"], additional_losses=additional)


class TrainingLosses(Losses):
    def __init__(
        self,
        backward: torch.Tensor | dict[str, torch.Tensor],
        additional_losses: dict[str, torch.Tensor] | None = None,
    ) -> None:
        super().__init__(additional_losses)
        self.backward: dict[str, torch.Tensor] = backward if isinstance(backward, dict) else {"backward": backward}

    def _named_scalars(self) -> dict[str, torch.Tensor]:
        out = super()._named_scalars()
        out.update(self.backward)
        return out

    def _flat_items(self) -> dict[str, torch.Tensor]:
        items = super()._flat_items()
        items.update({f"backward::{k}": v for k, v in self.backward.items()})
        return items

    @staticmethod
    def aggregate(loss_meter: LossMeter[TrainingLosses]) -> TrainingLosses:
        sums = loss_meter._reduced()
        additional = {k.split("::", 1)[1]: v for k, v in sums.items() if k.startswith("additional::")}
        backward = {k.split("::", 1)[1]: v for k, v in sums.items() if k.startswith("backward::")}
        if set(backward.keys()) == {"backward"}:
            return TrainingLosses(backward=backward["backward"], additional_losses=additional)
        return TrainingLosses(backward=backward, additional_losses=additional)


class LossMeterType(Enum):
    AVERAGE = "AVERAGE"
    ACCUMULATION = "ACCUMULATION"


LossesType = TypeVar("LossesType", bound=Losses)


class LossMeter(Generic[LossesType]):
    def __init__(self, loss_meter_type: LossMeterType, losses_type: type[LossesType]) -> None:
        self.loss_meter_type = loss_meter_type
        self.losses_type = losses_type
        self._sums: dict[str, torch.Tensor] = {}
        self._seen: dict[str, int] = {}  # per key: in how many steps of this window it was reported
        self.count = 0

    def update(self, losses: LossesType) -> None:
        self.accumulate(losses)
        self.mark_step(losses=losses)

    def accumulate(self, losses: LossesType) -> None:
        """Device-side part of ``update`` (safe inside CUDA-graph capture once every key has been seen)."""
        for key, value in losses._flat_items().items():
            value = value.detach()
            acc = self._sums.get(key)
            if acc is None:
                self._sums[key] = value.to(torch.float32).clone().reshape(())
            else:
                acc.add_(value.reshape(()))

    def mark_step(self, n: int = 1, losses: LossesType | None = None) -> None:
        """Host-side part of ``update``: count ``n`` steps (the accumulation itself may have been replayed by a captured
        graph).  ``losses`` names the keys those steps reported -- a key that comes and goes (an optional penalty, an
        ensemble member) is averaged over the steps that had it, like the reference's list of per-step dictionaries;
        without it every known key counts."""
        self.count += n
        for key in (losses._flat_items() if losses is not None else self._sums):
            self._seen[key] = self._seen.get(key, 0) + n

    def clear(self) -> None:
        # zero in place: accumulators referenced by captured graphs must keep their addresses.
        for acc in self._sums.values():
            acc.zero_()
        self._seen = {}
        self.count = 0

    def reset(self) -> None:
        self._sums, self._seen = {}, {}
        self.count = 0

    def _reduced(self) -> dict[str, torch.Tensor]:
        assert self.count > 0, "Cannot compute the aggregate of an empty loss meter"
        reported = {key: total for key, total in self._sums.items() if self._seen.get(key, 0) > 0}
        if self.loss_meter_type == LossMeterType.AVERAGE:
            return {key: total / self._seen[key] for key, total in reported.items()}
        return {key: total.clone() for key, total in reported.items()}

    def compute(self) -> LossesType:
        return self.losses_type.aggregate(self)  # type: ignore[return-value]

    @staticmethod
    def aggregate_losses_dict(
        loss_list: list[dict[str, torch.Tensor]], loss_meter_type: LossMeterType
    ) -> dict[str, torch.Tensor]:
        totals: dict[str, torch.Tensor] = {}
        present: dict[str, int] = {}
        for loss_dict in loss_list:
            for key, loss in loss_dict.items():
                totals[key] = totals[key] + loss if key in totals else loss.clone()
                present[key] = present.get(key, 0) + 1
        if loss_meter_type == LossMeterType.AVERAGE:  # over the dictionaries that carry the key
            return {key: total / present[key] for key, total in totals.items()}
        return totals

    # --- state (pickled by the client state checkpointer) -------------------------------------------------
    def __getstate__(self) -> dict:
        state = self.__dict__.copy()
        state["_sums"] = {k: v.detach().cpu() for k, v in self._sums.items()}
        return state

    def __setstate__(self, state: dict) -> None:
        self.__dict__.update(state)
        if "_seen" not in state:  # snapshots written before per-key step counts existed: every key was seen every step
            self._seen = dict.fromkeys(self._sums, self.count)
