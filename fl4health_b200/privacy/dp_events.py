"""Plain-data descriptions of DP mechanisms (the subset of ``dp_accounting.dp_event`` the reference builds in
``privacy/moments_accountant.py:26-61,113-140``).  The in-house accountant works on RDP curves directly; these events
exist so code written against the reference's ``SamplingStrategy.get_dp_event`` keeps working, and so a mechanism can be
logged / serialised.  ``to_dp_accounting`` converts to the real library's events when it is installed."""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Any


class NeighborRel(Enum):
    ADD_OR_REMOVE_ONE = 1
    REPLACE_ONE = 2
    REPLACE_SPECIAL = 3


@dataclass(frozen=True)
class DpEvent:
    def to_dp_accounting(self) -> Any:
        raise NotImplementedError


@dataclass(frozen=True)
class GaussianDpEvent(DpEvent):
    noise_multiplier: float

    def to_dp_accounting(self) -> Any:
        from dp_accounting import dp_event  # type: ignore[import-not-found]

        return dp_event.GaussianDpEvent(self.noise_multiplier)


@dataclass(frozen=True)
class PoissonSampledDpEvent(DpEvent):
    sampling_probability: float
    event: DpEvent

    def to_dp_accounting(self) -> Any:
        from dp_accounting import dp_event  # type: ignore[import-not-found]

        return dp_event.PoissonSampledDpEvent(self.sampling_probability, self.event.to_dp_accounting())


@dataclass(frozen=True)
class SampledWithoutReplacementDpEvent(DpEvent):
    source_dataset_size: int
    sample_size: int
    event: DpEvent

    def to_dp_accounting(self) -> Any:
        from dp_accounting import dp_event  # type: ignore[import-not-found]

        return dp_event.SampledWithoutReplacementDpEvent(self.source_dataset_size, self.sample_size, self.event.to_dp_accounting())


@dataclass(frozen=True)
class SelfComposedDpEvent(DpEvent):
    event: DpEvent
    count: int

    def to_dp_accounting(self) -> Any:
        from dp_accounting import dp_event  # type: ignore[import-not-found]

        return dp_event.SelfComposedDpEvent(self.event.to_dp_accounting(), self.count)
