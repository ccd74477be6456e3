"""``make_it_personal``: turn any ``FlexibleClient`` class into its Ditto / MR-MTL personalised variant (parity:
``fl4health/mixins/personalized/__init__.py:8-44``)."""

from __future__ import annotations

from enum import Enum

from fl4health_b200.clients.flexible.base import FlexibleClient
from fl4health_b200.mixins.personalized.ditto import DittoPersonalizedMixin
from fl4health_b200.mixins.personalized.mr_mtl import MrMtlPersonalizedMixin


class PersonalizedMode(Enum):
    DITTO = "ditto"
    MR_MTL = "mr_mtl"


def make_it_personal(client_base_type: type[FlexibleClient], mode: PersonalizedMode) -> type[FlexibleClient]:
    if mode == PersonalizedMode.DITTO:
        mixin, prefix = DittoPersonalizedMixin, "Ditto"
    elif mode == PersonalizedMode.MR_MTL:
        mixin, prefix = MrMtlPersonalizedMixin, "MrMtl"
    else:
        raise ValueError("Unrecognized personalized mode.")
    return type(f"{prefix}{client_base_type.__name__}", (mixin, client_base_type), {"_dynamically_created": True})


__all__ = ["DittoPersonalizedMixin", "MrMtlPersonalizedMixin", "PersonalizedMode", "make_it_personal"]
