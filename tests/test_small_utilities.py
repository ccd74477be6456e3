"""Small utilities that had no direct test: nnU-Net helpers that do not need nnU-Net (parity:
``tests/utils/nnnunet_utils_test.py``) and the model <-> DP-engine glue (``tests/utils/privacy_utilities_test.py``)."""

from __future__ import annotations

import logging
import os
import signal
import sys
import types

import pytest
import torch
from torch import nn

from fl4health_b200.privacy.dp_engine import GradSampleModule
from fl4health_b200.utils.nnunet_utils import (
    Module2LossWrapper,
    StreamToLogger,
    reload_modules,
    set_nnunet_env,
    use_default_signal_handlers,
)
from fl4health_b200.utils.privacy_utilities import (
    convert_model_to_opacus_model,
    map_model_to_opacus_model,
    privacy_validate_and_fix_modules,
)


def test_stream_to_logger_forwards_complete_lines(caplog: pytest.LogCaptureFixture) -> None:
    logger = logging.getLogger("fl4h.test.stream")
    stream = StreamToLogger(logger, logging.WARNING)
    with caplog.at_level(logging.DEBUG, logger="fl4h.test.stream"):
        written = stream.write("first line\nsecond line   \n")
        print("printed through the stream", file=stream)
        stream.flush()
    assert written == len("first line\nsecond line   \n")
    messages = [record.getMessage() for record in caplog.records if record.name == "fl4h.test.stream"]
    assert messages == ["first line", "second line", "printed through the stream"]
    assert {record.levelno for record in caplog.records if record.name == "fl4h.test.stream"} == {logging.WARNING}


def test_default_signal_handlers_inside_and_restored_after() -> None:
    def custom(signum, frame) -> None:  # noqa: ANN001
        raise AssertionError("must not be installed while the wrapped function runs")

    before = (signal.getsignal(signal.SIGINT), signal.getsignal(signal.SIGTERM))
    signal.signal(signal.SIGINT, custom)
    signal.signal(signal.SIGTERM, custom)
    try:
        seen = {}

        @use_default_signal_handlers
        def body(value: int) -> int:
            seen["int"], seen["term"] = signal.getsignal(signal.SIGINT), signal.getsignal(signal.SIGTERM)
            return value + 1

        assert body(1) == 2
        assert seen["int"] is signal.default_int_handler and seen["term"] == signal.SIG_DFL
        assert signal.getsignal(signal.SIGINT) is custom and signal.getsignal(signal.SIGTERM) is custom

        @use_default_signal_handlers
        def failing() -> None:
            raise KeyError("boom")

        with pytest.raises(KeyError):
            failing()
        assert signal.getsignal(signal.SIGINT) is custom  # restored on the error path too
    finally:
        signal.signal(signal.SIGINT, before[0])
        signal.signal(signal.SIGTERM, before[1])


def test_set_nnunet_env_and_module_reload(monkeypatch: pytest.MonkeyPatch) -> None:
    monkeypatch.delenv("nnUNet_raw", raising=False)
    set_nnunet_env(nnUNet_raw="/data/raw", nnUNet_results=7)
    assert os.environ["nnUNet_raw"] == "/data/raw" and os.environ["nnUNet_results"] == "7"
    monkeypatch.delenv("nnUNet_results", raising=False)

    # a module that reads the variable at import time sees the new value only after a reload
    package = types.ModuleType("fl4h_fake_nnunet")
    package.__path__ = []  # type: ignore[attr-defined]
    monkeypatch.setitem(sys.modules, "fl4h_fake_nnunet", package)
    reload_modules(["fl4h_fake_nnunet"])  # not reloadable (no spec): logged, not raised
    reload_modules(["a_package_that_was_never_imported"])


def test_module_to_loss_wrapper_is_a_loss_and_delegates() -> None:
    wrapped = Module2LossWrapper(nn.L1Loss(reduction="sum"))
    assert isinstance(wrapped, torch.nn.modules.loss._Loss)
    pred, target = torch.tensor([1.0, 2.0, 4.0]), torch.tensor([1.5, 2.0, 2.0])
    assert float(wrapped(pred, target)) == pytest.approx(2.5)


class _WithBatchNorm(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv2d(3, 8, 3, padding=1)
        self.norm = nn.BatchNorm2d(8)
        self.head = nn.Linear(8, 2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(torch.relu(self.norm(self.conv(x))).mean(dim=(2, 3)))


def test_privacy_validation_replaces_batchnorm_and_reports_reinitialisation() -> None:
    clean = nn.Sequential(nn.Linear(4, 4), nn.ReLU(), nn.Linear(4, 2))
    same, reinitialize = privacy_validate_and_fix_modules(clean)
    assert same is clean and reinitialize is False

    fixed, reinitialize = privacy_validate_and_fix_modules(_WithBatchNorm())
    assert reinitialize is True
    assert not any(isinstance(m, nn.modules.batchnorm._BatchNorm) for m in fixed.modules())
    assert any(isinstance(m, nn.GroupNorm) for m in fixed.modules())
    assert fixed(torch.randn(5, 3, 6, 6)).shape == (5, 2)  # still a working model


def test_conversion_wraps_once_and_mapping_fixes_then_wraps() -> None:
    wrapped = convert_model_to_opacus_model(nn.Linear(3, 2))
    assert isinstance(wrapped, GradSampleModule)
    assert convert_model_to_opacus_model(wrapped) is wrapped  # already wrapped: returned as is

    mapped = map_model_to_opacus_model(_WithBatchNorm())
    assert isinstance(mapped, GradSampleModule)
    assert not any(isinstance(m, nn.modules.batchnorm._BatchNorm) for m in mapped.modules())
    # per-sample gradients flow through the fixed model
    out = mapped(torch.randn(4, 3, 6, 6))
    out.sum().backward()
    head_weight = next(p for n, p in mapped.named_parameters() if n.endswith("head.weight"))
    assert head_weight.grad_sample.shape == (4, 2, 8)
