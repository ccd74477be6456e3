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


def test_reference_style_config_keys_are_understood() -> None:
    from examples.common import translate_reference_keys

    config = translate_reference_keys({
        "n_server_rounds": 15, "adapt_proximal_weight": True, "initial_proximal_weight": 0.0, "proximal_weight_delta": 0.1,
        "proximal_weight_patience": 5, "n_clients": 3, "local_epochs": 1, "local_steps": None, "batch_size": 128, "beta": 0.5,
        "initial_loss_weight": 0.3,  # a key given under both names keeps the native one
    })
    assert config["adapt_loss_weight"] is True and config["loss_weight_delta"] == 0.1 and config["loss_weight_patience"] == 5
    assert config["initial_loss_weight"] == 0.3 and config["heterogeneity_beta"] == 0.5
    assert "local_steps" not in config and config["local_epochs"] == 1 and "initial_proximal_weight" not in config


REFERENCE_EXAMPLES = "/root/reference/examples"


@pytest.mark.parametrize("scenario", ["fedprox_example", "apfl_example", "fedper_example", "feddg_ga_example", "moon_example"])
def test_scenario_runs_from_the_references_own_config_file(scenario: str, tmp_path, monkeypatch) -> None:
    """A YAML written for the reference (its key names, epochs instead of steps, its client count) drives the scenario."""
    import os

    import yaml

    source = os.path.join(REFERENCE_EXAMPLES, scenario, "config.yaml")
    if not os.path.exists(source):
        pytest.skip("reference examples are not on this machine")
    monkeypatch.chdir(tmp_path)
    config = yaml.safe_load(open(source).read())
    config.update({"samples_per_client": 48, "val_samples_per_client": 16, "batch_size": 16, "data_dir": str(tmp_path / "no_data_here")})
    path = tmp_path / "reference_style.yaml"
    path.write_text(yaml.safe_dump(config))
    summary = main([scenario, "--rounds", "1", "--device", "cpu", "--config", str(path)])
    assert summary["scenario"] == scenario and len(summary["losses"]) == 1 and summary["losses"][0][1] == summary["losses"][0][1]
