"""ctypes loader for ``libfl4h_ops.so`` + kernel-launch accounting."""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_LIB: ctypes.CDLL | None = None
_LOAD_ERROR: str | None = None
LIB_PATH = Path(__file__).resolve().parent / "libfl4h_ops.so"

# number of kernels of OUR library launched (host-side count; graph replays add their node counts explicitly)
_launch_count = 0


def count_launches(n: int = 1) -> None:
    global _launch_count
    _launch_count += n


def launch_count() -> int:
    return _launch_count


def reset_launch_count() -> None:
    global _launch_count
    _launch_count = 0


def load(required: bool = False) -> ctypes.CDLL | None:
    """Load the kernel library.  On a machine with a GPU a missing library is a hard error (no silent fallback)."""
    global _LIB, _LOAD_ERROR
    if _LIB is not None:
        return _LIB
    if _LOAD_ERROR is not None and not required:
        return None
    try:
        if not LIB_PATH.exists():
            if os.environ.get("FL4H_NO_AUTOBUILD") == "1":
                raise FileNotFoundError(str(LIB_PATH))
            from fl4health_b200.ops import build as _build

            _build.build()
        _LIB = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
    except Exception as exc:  # noqa: BLE001
        _LOAD_ERROR = f"{type(exc).__name__}: {exc}"
        if required or torch.cuda.is_available():
            raise RuntimeError(
                f"fl4health_b200 kernel library could not be loaded ({_LOAD_ERROR}); run "
                "`python -m fl4health_b200.ops.build`"
            ) from exc
        return None
    return _LIB


def available() -> bool:
    """True when kernels can actually run: library present AND a CUDA device is visible."""
    return torch.cuda.is_available() and load() is not None


def stream_ptr(device: torch.device | None = None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def check(err: int, what: str) -> None:
    if err != 0:
        raise RuntimeError(f"{what} failed with cudaError {err}")
