"""``make_it_personal``: turn any ``FlexibleClient`` class into its Ditto / MR-MTL personalised variant (parity:
``fl4health/mixins/personalized/__init__.py:8-44``)."""

from __future__ import annotations

from enum import Enum

from fl4health_b200.clients.flexible.base import FlexibleClient
from fl4health_b200.mixins.core_protocols import DittoPersonalizedProtocol, MrMtlPersonalizedProtocol
from fl4health_b200.mixins.personalized.ditto import DittoPersonalizedMixin
from fl4health_b200.mixins.personalized.mr_mtl import MrMtlPersonalizedMixin


class PersonalizedMode(Enum):
    DITTO = "ditto"
    MR_MTL = "mr_mtl"


# mode -> (mixin, class-name prefix); extend it to register further personalisation schemes
PersonalizedMixinRegistry: dict[PersonalizedMode, type] = {
    PersonalizedMode.DITTO: DittoPersonalizedMixin,
    PersonalizedMode.MR_MTL: MrMtlPersonalizedMixin,
}
_CLASS_PREFIX = {PersonalizedMode.DITTO: "Ditto", PersonalizedMode.MR_MTL: "MrMtl"}


def make_it_personal(client_base_type: type[FlexibleClient], mode: PersonalizedMode) -> type[FlexibleClient]:
    mixin = PersonalizedMixinRegistry.get(mode)
    if mixin is None:
        raise ValueError("Unrecognized personalized mode.")
    prefix = _CLASS_PREFIX.get(mode, str(getattr(mode, "value", mode)))
    return type(f"{prefix}{client_base_type.__name__}", (mixin, client_base_type), {"_dynamically_created": True})


__all__ = [
    "DittoPersonalizedMixin", "DittoPersonalizedProtocol", "MrMtlPersonalizedMixin", "MrMtlPersonalizedProtocol",
    "PersonalizedMixinRegistry", "PersonalizedMode", "make_it_personal",
]
