"""In-tree build of the sm_100a kernel library.

``python -m fl4health_b200.ops.build`` compiles every ``csrc/*.cu`` into ONE shared object
``fl4health_b200/ops/libfl4h_ops.so`` with plain ``nvcc`` (cross-compiles without a GPU).  The library exposes a C
ABI consumed through ``ctypes`` (see ``_lib.py``): no torch headers are involved, so the build takes seconds and the
``.so`` has no libtorch ABI coupling.  The ``.so`` is git-ignored but travels to GPU boxes with the working tree.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

OPS_DIR = Path(__file__).resolve().parent
CSRC = OPS_DIR / "csrc"
LIB_PATH = OPS_DIR / "libfl4h_ops.so"
STAMP_PATH = OPS_DIR / ".libfl4h_ops.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
    "-Xcompiler", "-fPIC", "-cudart", "shared",
    "--expt-relaxed-constexpr",
]


def _find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for src in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))):
        h.update(src.name.encode())
        h.update(src.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP_PATH.exists() and STAMP_PATH.read_text().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and is_current():
        return LIB_PATH
    nvcc = _find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found; cannot build fl4health_b200 kernels")
    objs = []
    build_dir = OPS_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in sources():
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{out}")
        if verbose and out:
            print(out)
    # no -lcuda: driver entry points are resolved at run time (cudaGetDriverEntryPoint) so the library loads on
    # driver-less build machines
    link = [nvcc, "-shared", "-cudart", "shared", "-o", str(LIB_PATH), *map(str, objs)]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    STAMP_PATH.write_text(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(f"built {path}")
