"""Method clients on a real GPU through the engine: every listed scenario runs with the arena, the fused flat optimizer
and CUDA-graph capture of its (overridden) train / val step enabled, for three rounds on synthetic shards.  The CPU
suite checks the methods' semantics; this checks that their hooks survive capture (or fall back cleanly) on a B200."""

import math

import pytest

from examples.run import main

pytestmark = pytest.mark.gpu

SCENARIOS = [
    "basic_example", "fedprox_example", "scaffold_example", "ditto_example", "mr_mtl_example", "moon_example", "apfl_example",
    "fedper_example", "fedrep_example", "fedbn_example", "fenda_example", "perfcl_example", "gpfl_example", "fedpm_example",
    "instance_level_dp_example", "client_level_dp_example", "flash_example", "fedopt_example", "ensemble_example",
    "sparse_tensor_partial_exchange_example",
]


@pytest.mark.parametrize("scenario", SCENARIOS)
def test_scenario_trains_on_gpu_with_graphs(scenario: str, tmp_path, monkeypatch) -> None:
    import yaml

    from examples.common import CONFIG_DIR

    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("FL4H_CUDA_GRAPHS", "1")
    monkeypatch.setenv("FL4H_LOG_LEVEL", "WARNING")
    config = yaml.safe_load((CONFIG_DIR / f"{scenario}.yaml").read_text()) or {}
    config.update({"samples_per_client": 256, "val_samples_per_client": 64, "local_steps": 8, "batch_size": 32,
                   "data_dir": str(tmp_path / "no_data_here")})
    if "local_epochs" in config:
        config["local_epochs"] = 1
    path = tmp_path / f"{scenario}.yaml"
    path.write_text(yaml.safe_dump(config))
    summary = main([scenario, "--rounds", "3", "--clients", "2", "--device", "cuda", "--config", str(path)])
    losses = [loss for _, loss in summary["losses"]]
    assert len(losses) == 3 and all(math.isfinite(loss) for loss in losses), losses
