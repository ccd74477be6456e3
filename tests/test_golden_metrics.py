"""End-to-end golden-metric tests: each scenario's ``JsonReporter`` output (server and every client, every round) must
reproduce the checked-in goldens to the reference's smoke-test tolerance (``tests/smoke_tests/run_smoke_test.py``:
``DEFAULT_TOLERANCE = 0.0005``).  A silent change in an aggregation rule, a loss term, a metric, the number of local
steps, the order of client sampling or the seeding shows up here, not just "ran without NaN".

Regenerate after an intended behaviour change with ``python tools/make_goldens.py [scenario ...]``.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path
from typing import Any

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))

from make_goldens import GOLDEN_DIR, GOLDEN_SCENARIOS, run_scenario  # noqa: E402

TOLERANCE = 5e-4


def _mismatches(golden: Any, actual: Any, where: str = "") -> list[str]:
    if isinstance(golden, dict):
        if not isinstance(actual, dict):
            return [f"{where}: expected a dict, found {type(actual).__name__}"]
        problems = [f"{where}/{key}: missing from the run" for key in golden.keys() - actual.keys()]
        problems += [f"{where}/{key}: not in the golden" for key in actual.keys() - golden.keys()]
        for key in golden.keys() & actual.keys():
            problems += _mismatches(golden[key], actual[key], f"{where}/{key}")
        return problems
    if isinstance(golden, int) and not isinstance(golden, bool) and isinstance(actual, int):
        return [] if golden == actual else [f"{where}: {actual} != {golden}"]
    close = abs(actual - golden) <= TOLERANCE * max(1.0, abs(golden))
    return [] if close else [f"{where}: {actual} vs golden {golden}"]


@pytest.mark.parametrize("scenario", GOLDEN_SCENARIOS)
def test_reports_match_goldens(scenario: str, tmp_path: Path) -> None:
    golden = json.loads((GOLDEN_DIR / f"{scenario}.json").read_text())
    assert set(golden) == {"server", "client_0", "client_1"} and set(golden["server"]["rounds"]) == {"1", "2"}
    actual = run_scenario(scenario, tmp_path)
    problems = _mismatches(golden, actual)
    assert not problems, "\n".join(problems[:20])


def test_the_comparison_has_teeth() -> None:
    golden = json.loads((GOLDEN_DIR / "basic_example.json").read_text())
    assert _mismatches(golden, golden) == []
    drifted = json.loads(json.dumps(golden))
    drifted["server"]["rounds"]["2"]["val - loss - aggregated"] *= 1.002  # a 0.2 % change in one aggregated loss
    drifted["client_1"]["rounds"]["1"]["fit_step"] += 1
    del drifted["client_0"]["rounds"]["2"]["eval_round_loss"]
    problems = _mismatches(golden, drifted)
    assert len(problems) == 3 and any("fit_step" in p for p in problems) and any("missing" in p for p in problems)
