"""Small helpers on the per-round critical path: cheap train/eval switch, logger fast-reject, batched scalar reads,
integer-state plumbing through Parameters."""

import logging

import torch
from torch import nn

from fl4health_b200.common import logger as fl_logger
from fl4health_b200.common.typing import NDArrays, ndarrays_to_parameters, parameters_to_ndarrays
from fl4health_b200.engine.modes import invalidate, set_training
from fl4health_b200.utils.losses import EvaluationLosses, TrainingLosses, read_scalars


def test_set_training_matches_module_train() -> None:
    model = nn.Sequential(nn.Conv2d(1, 2, 3), nn.BatchNorm2d(2), nn.Sequential(nn.Dropout(0.5), nn.Linear(2, 2)))
    set_training(model, False)
    assert not any(m.training for m in model.modules())
    set_training(model, True)
    assert all(m.training for m in model.modules())
    # a newly added sub-module is picked up after invalidate()
    model.add_module("extra", nn.Dropout(0.1))
    invalidate(model)
    set_training(model, False)
    assert not model.extra.training


def test_set_training_respects_overridden_train() -> None:
    class FrozenBn(nn.BatchNorm2d):
        def train(self, mode: bool = True):  # noqa: ANN202 - frozen-statistics recipe
            return super().train(False)

    model = nn.Sequential(nn.Conv2d(1, 2, 3), FrozenBn(2))
    set_training(model, True)
    assert model[0].training and not model[1].training  # fell back to nn.Module.train(): the override was honoured


def test_logger_fast_reject_and_user_handlers() -> None:
    # pytest's logging plugin hangs level-0 capture handlers on the logger: park them so the level logic is visible
    parked = [h for h in fl_logger.FLOWER_LOGGER.handlers if h is not fl_logger.console_handler]
    for handler in parked:
        fl_logger.FLOWER_LOGGER.removeHandler(handler)
    try:
        _check_logger_levels()
    finally:
        for handler in parked:
            fl_logger.FLOWER_LOGGER.addHandler(handler)
        fl_logger.update_console_handler(level=logging.INFO)


def _check_logger_levels() -> None:
    fl_logger.update_console_handler(level=logging.WARNING)
    assert fl_logger.FLOWER_LOGGER.level == logging.WARNING

    class Sink(logging.Handler):
        def __init__(self) -> None:
            super().__init__(level=logging.DEBUG)
            self.records: list[str] = []

        def emit(self, record: logging.LogRecord) -> None:
            self.records.append(record.getMessage())

    sink = Sink()
    fl_logger.log(logging.INFO, "dropped before a LogRecord exists")
    fl_logger.FLOWER_LOGGER.addHandler(sink)
    try:
        fl_logger.log(logging.INFO, "seen by the user handler")
        assert sink.records == ["seen by the user handler"]
    finally:
        fl_logger.FLOWER_LOGGER.removeHandler(sink)
        fl_logger.log(logging.INFO, "dropped again")  # re-syncs the level
        assert fl_logger.FLOWER_LOGGER.level == logging.WARNING


def test_read_scalars_and_loss_dicts() -> None:
    values = {"a": torch.tensor(1.5), "b": torch.tensor(2, dtype=torch.int64), "c": torch.tensor(0.25, dtype=torch.bfloat16)}
    assert read_scalars(values) == {"a": 1.5, "b": 2.0, "c": 0.25}
    train = TrainingLosses(torch.tensor(0.5), {"penalty": torch.tensor(0.125)})
    assert train.as_dict() == {"penalty": 0.125, "backward": 0.5}
    evaluation = EvaluationLosses(torch.tensor(2.0), {"extra": torch.tensor(4.0)})
    assert evaluation.as_dict() == {"extra": 4.0, "checkpoint": 2.0}


def test_int_flat_survives_parameters_round_trip() -> None:
    arrays = NDArrays([torch.zeros(3), torch.tensor(7)])
    arrays.int_flat = torch.tensor([7])
    back = parameters_to_ndarrays(ndarrays_to_parameters(arrays))
    assert back.int_flat is arrays.int_flat
    assert parameters_to_ndarrays(ndarrays_to_parameters([torch.zeros(2)])).int_flat is None


def test_optimizer_references_skip_zero_coefficients_on_infinite_parameters() -> None:
    """0 * inf = NaN: with zero weight decay / drift weight the terms are skipped (torch.optim semantics), so FedPM's
    +-inf scores (sigmoid_inverse of an aggregate of exactly 0 or 1) survive a local step."""
    from fl4health_b200.ops import flat as F

    w = torch.tensor([float("inf"), -float("inf"), 1.0, -2.0])
    hp = F.make_hyper_params("cpu")
    hp[F.HP_LR] = 0.5
    F.sgd_step_reference(w, torch.full((4,), 0.25), None, hp, anchor=torch.zeros(4))
    assert not torch.isnan(w).any() and torch.isinf(w[:2]).all() and torch.allclose(w[2:], torch.tensor([0.875, -2.125]))
    w = torch.tensor([float("inf"), 1.0, -2.0, 0.5])
    m, v = torch.zeros(4), torch.zeros(4)
    hp = F.make_hyper_params("cpu")
    hp[F.HP_LR], hp[F.HP_B1], hp[F.HP_B2], hp[F.HP_EPS] = 0.1, 0.9, 0.999, 1e-8
    F.adamw_step_reference(w, torch.full((4,), 0.25), m, v, hp, anchor=torch.zeros(4))
    assert not torch.isnan(w).any() and torch.isinf(w[0])
