"""Process-wide ``flwr`` logger."""

from __future__ import annotations

import logging
import os

LOGGER_NAME = "flwr"
FLOWER_LOGGER = logging.getLogger(LOGGER_NAME)
FLOWER_LOGGER.setLevel(getattr(logging, os.environ.get("FLWR_SHIM_LOG_LEVEL", "INFO")))

console_handler = logging.StreamHandler()
console_handler.setLevel(logging.DEBUG)
console_handler.setFormatter(logging.Formatter("%(levelname)s :      %(message)s"))
FLOWER_LOGGER.addHandler(console_handler)
FLOWER_LOGGER.propagate = False

log = FLOWER_LOGGER.log


def configure(identifier: str, filename: str | None = None, host: str | None = None) -> None:
    if filename:
        handler = logging.FileHandler(filename)
        handler.setFormatter(logging.Formatter(f"{identifier} | %(levelname)s %(asctime)s | %(message)s"))
        FLOWER_LOGGER.addHandler(handler)
