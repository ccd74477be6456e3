"""Wire layouts of the packers, the layer exchangers, the snapshotters and the reporters (mirrors the reference's
tests/parameter_exchange, tests/utils/snapshotter_test.py, tests/reporting)."""

import json
import sys
import types

import numpy as np
import pytest
import torch
from torch import nn

from fl4health_b200.common.typing import NDArrays
from fl4health_b200.parameter_exchange.layer_exchanger import FixedLayerExchanger, LayerExchangerWithExclusions
from fl4health_b200.parameter_exchange.packing_exchanger import FullParameterExchangerWithPacking
from fl4health_b200.parameter_exchange.parameter_packer import (
    ParameterPackerAdaptiveConstraint,
    ParameterPackerWithClippingBit,
    ParameterPackerWithControlVariates,
    ParameterPackerWithLayerNames,
    SparseCooParameterPacker,
)
from fl4health_b200.reporting import JsonReporter
from fl4health_b200.reporting.base_reporter import BaseReporter
from fl4health_b200.reporting.reports_manager import ReportsManager
from fl4health_b200.utils import snapshotter as snap


def _weights() -> NDArrays:
    return NDArrays([torch.arange(4.0).view(2, 2), torch.tensor([1.0, 2.0]), torch.tensor([[3.0]])])


def test_packer_wire_layouts() -> None:
    weights = _weights()
    variates = NDArrays([w * 10 for w in weights])
    packer = ParameterPackerWithControlVariates(len(weights))
    packed = packer.pack_parameters(weights, variates)
    assert len(packed) == 6  # weights ++ variates, split at size_of_model_params
    w2, v2 = packer.unpack_parameters(packed)
    assert all(torch.equal(a, b) for a, b in zip(w2, weights)) and all(torch.equal(a, b) for a, b in zip(v2, variates))

    for cls, extra in ((ParameterPackerWithClippingBit, 1.0), (ParameterPackerAdaptiveConstraint, 0.25)):
        packed = cls().pack_parameters(weights, extra)
        assert len(packed) == len(weights) + 1  # weights ++ [scalar]
        w2, value = cls().unpack_parameters(packed)
        assert value == extra and len(w2) == len(weights)

    names = ["conv.weight", "conv.bias", "fc.weight"]
    packed = ParameterPackerWithLayerNames().pack_parameters(weights, names)
    assert isinstance(packed[-1], np.ndarray) and packed[-1].dtype.kind in "US"  # names ride as a numpy string array
    w2, names2 = ParameterPackerWithLayerNames().unpack_parameters(packed)
    assert names2 == names and len(w2) == 3

    dense = torch.tensor([[0.0, 2.0, 0.0], [0.0, 0.0, 5.0]])
    values, indices, shape = SparseCooParameterPacker.extract_coo_info_from_dense(dense)
    assert values.tolist() == [2.0, 5.0] and indices.tolist() == [[0, 1], [1, 2]] and shape.tolist() == [2, 3]
    packed = SparseCooParameterPacker().pack_parameters(NDArrays([values]), (NDArrays([indices]), NDArrays([shape]), ["layer"]))
    assert len(packed) == 4  # values ++ indices ++ shapes ++ [names]
    v2, (i2, s2, n2) = SparseCooParameterPacker().unpack_parameters(packed)
    assert torch.equal(v2[0], values) and torch.equal(i2[0], indices) and n2 == ["layer"]

    exchanger = FullParameterExchangerWithPacking(ParameterPackerAdaptiveConstraint())
    model = nn.Linear(2, 2)
    packed = exchanger.pack_parameters(exchanger.push_parameters(model), 0.5)
    pushed, mu = exchanger.unpack_parameters(packed)
    assert mu == 0.5 and len(pushed) == 2


def test_fixed_and_exclusion_layer_exchangers() -> None:
    torch.manual_seed(0)
    source = nn.Sequential(nn.Conv2d(1, 2, 3), nn.BatchNorm2d(2), nn.Flatten(), nn.Linear(8, 2))
    target = nn.Sequential(nn.Conv2d(1, 2, 3), nn.BatchNorm2d(2), nn.Flatten(), nn.Linear(8, 2))
    source[1].running_mean.fill_(0.7)
    fixed = FixedLayerExchanger(["0.weight", "0.bias"])
    payload = fixed.push_parameters(source)
    assert len(payload) == 2
    before_fc = target[3].weight.detach().clone()
    fixed.pull_parameters(payload, target)
    assert torch.equal(target[0].weight, source[0].weight) and torch.equal(target[3].weight, before_fc)

    fedbn = LayerExchangerWithExclusions(source, {nn.BatchNorm2d})
    names = fedbn.get_layers_to_transfer(source)
    assert all(not n.startswith("1.") for n in names) and "3.weight" in names  # every BN key (buffers included) stays local
    payload = fedbn.push_parameters(source)
    fedbn.pull_parameters(payload, target)
    assert torch.equal(target[3].weight, source[3].weight)
    assert not torch.equal(target[1].running_mean, source[1].running_mean)  # BN statistics were not exchanged


def test_snapshotters_round_trip() -> None:
    model = nn.Linear(3, 2)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=0.5)
    model(torch.randn(4, 3)).sum().backward()
    optimizer.step()
    scheduler.step()
    saved_opt = snap.OptimizerSnapshotter().save_attribute({"global": optimizer})
    saved_sched = snap.LRSchedulerSnapshotter().save_attribute({"global": scheduler})
    saved_model = snap.TorchModuleSnapshotter().save_attribute({"model": model})
    assert set(saved_opt["global"].keys()) == set(optimizer.state_dict()["state"].keys())  # state only, no param groups

    fresh = nn.Linear(3, 2)
    fresh_opt = torch.optim.SGD(fresh.parameters(), lr=0.1, momentum=0.9)
    fresh_sched = torch.optim.lr_scheduler.StepLR(fresh_opt, step_size=1, gamma=0.5)
    fresh(torch.randn(4, 3)).sum().backward()
    fresh_opt.step()
    snap.TorchModuleSnapshotter().load_attribute(saved_model, {"model": fresh})
    snap.OptimizerSnapshotter().load_attribute(saved_opt, {"global": fresh_opt})
    snap.LRSchedulerSnapshotter().load_attribute(saved_sched, {"global": fresh_sched})
    assert torch.equal(fresh.weight, model.weight)
    momentum = [s["momentum_buffer"] for s in optimizer.state.values()]
    restored = [s["momentum_buffer"] for s in fresh_opt.state.values()]
    assert all(torch.equal(a, b) for a, b in zip(momentum, restored))
    assert fresh_sched.last_epoch == scheduler.last_epoch

    for cls, value in ((snap.SingletonSnapshotter, 7), (snap.StringSnapshotter, "name"), (snap.BytesSnapshotter, b"plans"),
                       (snap.SerializableObjectSnapshotter, {"k": [1, 2]})):
        holder = {"attr": None}
        cls().load_attribute(cls().save_attribute({"attr": value}), holder)
        assert holder["attr"] == value


def test_json_reporter_and_reports_manager(tmp_path) -> None:
    reporter = JsonReporter(run_id="run7", output_folder=tmp_path)

    class Recorder(BaseReporter):
        def __init__(self) -> None:
            self.calls: list = []

        def initialize(self, **kwargs) -> None:
            self.calls.append(("init", kwargs))

        def report(self, data, round=None, epoch=None, step=None) -> None:  # noqa: A002
            self.calls.append((data, round, epoch, step))

        def shutdown(self) -> None:
            self.calls.append("shutdown")

    recorder = Recorder()
    manager = ReportsManager([reporter, recorder])
    manager.initialize(id="client_a")
    manager.report({"host_type": "client"})
    manager.report({"fit_round_losses": {"backward": torch.tensor(0.5)}, "round": 1}, 1)
    manager.report({"fit_step_losses": {"backward": 0.4}}, 1, None, 3)  # per-step data: ignored by the JSON reporter
    manager.report({"eval_round_loss": np.float32(0.25)}, 1)
    manager.shutdown()
    written = json.loads((tmp_path / "run7.json").read_text())
    assert written == {"host_type": "client", "rounds": {"1": {"fit_round_losses": {"backward": 0.5}, "round": 1, "eval_round_loss": 0.25}}}
    assert recorder.calls[0] == ("init", {"id": "client_a"}) and recorder.calls[-1] == "shutdown" and len(recorder.calls) == 6


def test_wandb_reporter_with_stub_module(monkeypatch) -> None:
    logged: list = []

    class Run:
        _run_id = "abc"

        def define_metric(self, name: str) -> None:
            logged.append(("define", name))

        def log(self, data: dict) -> None:
            logged.append(("log", data))

        def finish(self) -> None:
            logged.append("finish")

    stub = types.ModuleType("wandb")
    stub.init = lambda **kwargs: (logged.append(("init", kwargs)), Run())[1]  # type: ignore[attr-defined]
    monkeypatch.setitem(sys.modules, "wandb", stub)
    from fl4health_b200.reporting import WandBReporter, WandBStepType

    reporter = WandBReporter(WandBStepType.ROUND, project="p", tags=["t"])
    reporter.initialize(id="client_a", name="run")
    reporter.report({"round": 1, "fit_round_metrics": {"train - prediction - accuracy": 0.5}}, 1)
    reporter.report({"fit_step_losses": {"backward": 0.1}}, 1, None, 4)  # step data is dropped for ROUND granularity
    reporter.shutdown()
    init = next(item for item in logged if item[0] == "init")
    assert init[1]["project"] == "p" and init[1]["id"] == "client_a" and init[1]["name"] == "run"
    logs = [item[1] for item in logged if item[0] == "log"]
    assert logs == [{"round": 1, "train - prediction - accuracy": 0.5}] and logged[-1] == "finish"
    assert reporter.get_wandb_timestep(3, 2, 1) == 3
    assert WandBReporter("step").get_wandb_timestep(3, 2, 1) == 1
