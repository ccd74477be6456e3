"""EarlyStopper (parity: fl4health/utils/early_stopper.py:14-98): validates every ``interval_steps`` steps, snapshots
the client's state at each new best, restores the best state and stops once patience runs out."""

from pathlib import Path

import torch

from fl4health_b200.utils.early_stopper import EarlyStopper
from fl4health_b200.utils.random import set_all_random_seeds
from tests.helpers import fit_config_fn, make_clients


def _client(tmp_path: Path, patience, interval_steps: int, lr: float):
    set_all_random_seeds(5)
    client = make_clients(1, lr=lr)[0]
    config = fit_config_fn(local_steps=20)(1)
    client.setup_client(config)
    client.early_stopper = EarlyStopper(client, patience=patience, interval_steps=interval_steps, train_loop_checkpoint_dir=tmp_path)
    return client, config


def test_early_stopper_stops_and_restores_best_state(tmp_path: Path) -> None:
    # an absurd learning rate makes the validation loss blow up right after the first check: patience 1 must stop the
    # loop at the second check and put the step-0 weights back
    client, config = _client(tmp_path, patience=1, interval_steps=2, lr=50.0)
    calls = []
    original = client.early_stopper.should_stop

    def spy(step: int) -> bool:
        stop = original(step)
        if step % 2 == 0:
            calls.append((step, stop, {k: v.detach().clone() for k, v in client.model.state_dict().items()}))
        return stop

    client.early_stopper.should_stop = spy
    client.train_by_steps(20, current_round=1)
    assert (tmp_path / f"temp_{client.client_name}.pt").exists()
    stops = [step for step, stop, _ in calls if stop]
    assert stops and stops[0] < 19, calls  # stopped well before the 20 requested steps
    best_state = calls[0][2]  # state at the first check (the best: everything after diverged)
    for key, value in client.model.state_dict().items():
        if value.is_floating_point():
            assert torch.allclose(value, best_state[key]), key
    assert client.total_steps < 20


def test_early_stopper_without_patience_never_stops_but_tracks_best(tmp_path: Path) -> None:
    client, config = _client(tmp_path, patience=None, interval_steps=5, lr=0.01)
    client.train_by_steps(20, current_round=1)
    assert client.total_steps == 20
    assert client.early_stopper.best_score is not None
    # the best local state is reloaded into the model before parameters are shipped (patience=None semantics)
    client.early_stopper.load_snapshot(["model"])


def test_early_stopper_snapshot_dir_alias(tmp_path: Path) -> None:
    client, _ = _client(tmp_path, patience=2, interval_steps=3, lr=0.01)
    stopper = EarlyStopper(client, patience=2, interval_steps=3, snapshot_dir=tmp_path / "a")
    assert Path(stopper.state_checkpointer.checkpoint_dir) == tmp_path / "a"
