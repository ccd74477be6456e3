"""Placeholders, see ``dp_accounting/__init__.py``."""
from enum import Enum

from dp_accounting import _Unavailable


class NeighborRel(Enum):
    ADD_OR_REMOVE_ONE = 1
    REPLACE_ONE = 2
    REPLACE_SPECIAL = 3


class RdpAccountant(_Unavailable): ...
