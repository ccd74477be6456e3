"""Smoke tests: every example scenario builds and runs for two rounds on tiny synthetic shards (mirrors the
reference's examples smoke tests, tests/smoke_tests)."""

import pytest

from examples.run import main
from examples.scenarios import SCENARIOS

ONE_ROUND = {"fedpca_example"}
ONE_SHOT = {"federated_eval_example", "model_merge_example"}  # a single evaluation / merge pass: metrics, no per-round losses


@pytest.mark.parametrize("scenario", sorted(SCENARIOS))
def test_example_scenario_runs(scenario: str, tmp_path, monkeypatch) -> None:
    monkeypatch.chdir(tmp_path)  # examples write outputs relative to the working directory
    rounds = 1 if scenario in ONE_ROUND else 2
    summary = main([scenario, "--rounds", str(rounds), "--clients", "2", "--device", "cpu", "--config", _tiny_config(tmp_path, scenario)])
    assert summary["scenario"] == scenario
    if scenario in ONE_SHOT:
        assert summary["metrics"] and all(v == v for v in summary["metrics"].values())
        return
    assert len(summary["losses"]) == rounds
    assert all(loss == loss for _, loss in summary["losses"])  # no NaNs


def _tiny_config(tmp_path, scenario: str) -> str:
    import yaml

    from examples.common import CONFIG_DIR

    config = yaml.safe_load((CONFIG_DIR / f"{scenario}.yaml").read_text()) or {}
    config.update({"samples_per_client": 64, "val_samples_per_client": 32, "local_steps": 2, "batch_size": 16,
                   "data_dir": str(tmp_path / "no_data_here")})
    if "local_epochs" in config:
        config["local_epochs"] = 1
    path = tmp_path / f"{scenario}.yaml"
    path.write_text(yaml.safe_dump(config))
    return str(path)
