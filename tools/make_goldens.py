"""(Re)generate the golden per-round metrics that ``tests/test_golden_metrics.py`` diffs against.

Every scenario is run on CPU for two rounds on the tiny synthetic shards of the smoke tests, with a ``JsonReporter`` on
the server and on every client (the reference's smoke tests do the same against real datasets,
``tests/smoke_tests/run_smoke_test.py:733-783``); the numeric, non-timing entries of the reports are the golden.

    python tools/make_goldens.py                 # all scenarios listed in GOLDEN_SCENARIOS
    python tools/make_goldens.py ditto_example   # just one
"""

from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path
from typing import Any

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN_DIR = ROOT / "tests" / "goldens"
GOLDEN_SCENARIOS = [
    "basic_example", "fedopt_example", "fedprox_example", "scaffold_example", "ditto_example", "mr_mtl_example", "apfl_example",
    "moon_example", "fenda_example", "perfcl_example", "fedper_example", "fedrep_example", "fedbn_example", "gpfl_example",
    "flash_example", "feddg_ga_example", "ensemble_example", "dynamic_layer_exchange_example", "sparse_tensor_partial_exchange_example",
    "fedpm_example", "instance_level_dp_example", "client_level_dp_example",
]
_TIMING = ("start", "end", "elapsed", "seconds", "initialized", "shutdown", "time")
TINY = {"samples_per_client": 64, "val_samples_per_client": 32, "local_steps": 2, "batch_size": 16}


def numeric_view(report: Any) -> Any:
    """The part of a reporter payload that must reproduce: numbers (and nested dicts of them) under non-timing keys."""
    if isinstance(report, dict):
        kept = {key: numeric_view(value) for key, value in report.items() if not any(word in key for word in _TIMING)}
        return {key: value for key, value in kept.items() if value is not None and value != {}}
    if isinstance(report, bool) or not isinstance(report, (int, float)):
        return None
    return report


def run_scenario(scenario: str, workdir: Path) -> dict[str, Any]:
    import os

    import yaml

    from examples.common import CONFIG_DIR
    from examples.run import main

    config = yaml.safe_load((CONFIG_DIR / f"{scenario}.yaml").read_text()) or {}
    config.update({**TINY, "data_dir": str(workdir / "no_data_here")})
    if "local_epochs" in config:
        config["local_epochs"] = 1
    config_path = workdir / f"{scenario}.yaml"
    config_path.write_text(yaml.safe_dump(config))
    metrics_dir = workdir / "metrics"
    previous = os.getcwd()
    os.chdir(workdir)  # examples write their artefacts relative to the working directory
    try:
        main([scenario, "--rounds", "2", "--clients", "2", "--device", "cpu", "--config", str(config_path), "--metrics-dir", str(metrics_dir)])
    finally:
        os.chdir(previous)
    return {path.stem: numeric_view(json.loads(path.read_text())) for path in sorted(metrics_dir.glob("*.json"))}


def main() -> None:
    import os

    os.environ.setdefault("FL4H_LOG_LEVEL", "ERROR")
    GOLDEN_DIR.mkdir(exist_ok=True)
    for scenario in sys.argv[1:] or GOLDEN_SCENARIOS:
        with tempfile.TemporaryDirectory() as first, tempfile.TemporaryDirectory() as second:
            golden, again = run_scenario(scenario, Path(first)), run_scenario(scenario, Path(second))
        if golden != again:
            print(f"!! {scenario}: two runs differ, not written")
            continue
        (GOLDEN_DIR / f"{scenario}.json").write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")
        print(f"ok {scenario}")


if __name__ == "__main__":
    main()
