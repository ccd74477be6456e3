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

pytestmark = pytest.mark.skipif(not (REFERENCE / "fl4health").is_dir(), reason="reference arm not installed (baseline/install_reference.sh)")


@pytest.mark.parametrize("script", SCRIPTS, ids=[s.stem for s in SCRIPTS])
def test_agrees_with_the_reference(script: Path) -> None:
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT / "baseline" / "stubs"), str(REFERENCE), str(ROOT)]),
               CUDA_VISIBLE_DEVICES="")
    run = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=1200, cwd="/tmp")
    assert run.returncode == 0, (run.stdout[-1500:], run.stderr[-3000:])
    last = [line for line in run.stdout.splitlines() if line.startswith("configs agree:")]
    assert last and int(last[-1].split(":")[1]) > 0, run.stdout[-1500:]
