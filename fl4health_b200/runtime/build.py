"""In-tree build of the native host runtime (``libfl4h_runtime.so``): plain C++17, no CUDA, so it also builds and runs
on CPU-only machines (the gloo test suite uses it).  ``python -m fl4health_b200.runtime.build``."""

from __future__ import annotations

import hashlib
import shutil
import subprocess
import sys
from pathlib import Path

RUNTIME_DIR = Path(__file__).resolve().parent
CSRC = RUNTIME_DIR / "csrc"
LIB_PATH = RUNTIME_DIR / "libfl4h_runtime.so"
STAMP_PATH = RUNTIME_DIR / ".libfl4h_runtime.stamp"
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-pthread"]


def _fingerprint() -> str:
    h = hashlib.sha256()
    for src in sorted(CSRC.glob("*.cpp")):
        h.update(src.name.encode())
        h.update(src.read_bytes())
    h.update(" ".join(CXX_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP_PATH.exists() and STAMP_PATH.read_text().strip() == _fingerprint()


def build(force: bool = False) -> Path:
    if not force and is_current():
        return LIB_PATH
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        raise RuntimeError("no C++ compiler found; cannot build the fl4health_b200 host runtime")
    cmd = [cxx, *CXX_FLAGS, "-o", str(LIB_PATH), *map(str, sorted(CSRC.glob("*.cpp"))), "-lrt"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"host runtime build failed:\n{res.stdout}\n{res.stderr}")
    STAMP_PATH.write_text(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
