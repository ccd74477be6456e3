"""Differential tests: our modules against the UNMODIFIED reference package (``baseline/_ref``, driven through the
stand-alone stubs in ``baseline/stubs``) on the same inputs.  Each script under ``tests/differential/`` runs in its own
interpreter (the reference and its stubs never enter the test process) and prints ``configs agree: N``."""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = ROOT / "baseline" / "_ref"
SCRIPTS = sorted((Path(__file__).parent / "differential").glob("check_*.py"))



def _ensure_reference_copy() -> None:
    """Same outcome as ``baseline/install_reference.sh`` / ``reference_arm.reference_available``: an unmodified copy of
    the pure-Python reference package (``baseline/_ref`` is git-ignored, so a fresh checkout does not have it)."""
    source = Path(os.environ.get("FL4H_REFERENCE_SRC", "/root/reference")) / "fl4health"
    if not (REFERENCE / "fl4health" / "__init__.py").exists() and (source / "__init__.py").exists():
        import shutil

        shutil.copytree(source, REFERENCE / "fl4health", ignore=shutil.ignore_patterns("__pycache__"), dirs_exist_ok=True)


_ensure_reference_copy()
pytestmark = pytest.mark.skipif(not (REFERENCE / "fl4health").is_dir(), reason="reference arm not installed (baseline/install_reference.sh)")


@pytest.mark.parametrize("script", SCRIPTS, ids=[s.stem for s in SCRIPTS])
def test_agrees_with_the_reference(script: Path) -> None:
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT / "baseline" / "stubs"), str(REFERENCE), str(ROOT)]),
               CUDA_VISIBLE_DEVICES="")
    run = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=1200, cwd="/tmp")
    assert run.returncode == 0, (run.stdout[-1500:], run.stderr[-3000:])
    last = [line for line in run.stdout.splitlines() if line.startswith("configs agree:")]
    assert last and int(last[-1].split(":")[1]) > 0, run.stdout[-1500:]
