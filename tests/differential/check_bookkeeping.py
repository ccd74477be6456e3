"""Bookkeeping components vs the reference: loss meters, checkpointers (decisions, file interchangeability), the JSON
reporter's output document, warm-starting from a pretrained model, the fixed-sample client manager's contract."""
import json
import tempfile
from pathlib import Path

import torch
from torch import nn

import fl4health.checkpointing.checkpointer as ref_ckpt
import fl4health.preprocessing.warmed_up_module as ref_warm
import fl4health.reporting.json_reporter as ref_report
import fl4health.utils.losses as ref_losses
import fl4health_b200.checkpointing.checkpointer as my_ckpt
import fl4health_b200.preprocessing.warmed_up_module as my_warm
import fl4health_b200.reporting.json_reporter as my_report
import fl4health_b200.utils.losses as my_losses

torch.manual_seed(51)
agreed = 0


def close(a: dict, b: dict) -> None:
    assert a.keys() == b.keys(), (a, b)
    assert all(abs(float(a[k]) - float(b[k])) < 1e-6 for k in a), (a, b)


# -- loss meters ------------------------------------------------------------------------------------------------------
for meter_type in ("AVERAGE", "ACCUMULATION"):
    theirs = ref_losses.LossMeter(getattr(ref_losses.LossMeterType, meter_type), ref_losses.TrainingLosses)
    ours = my_losses.LossMeter(getattr(my_losses.LossMeterType, meter_type), my_losses.TrainingLosses)
    for step in range(5):
        backward = torch.rand(()) if step % 2 else {"backward": torch.rand(()), "second": torch.rand(())}
        extra = {"penalty": torch.rand(())} if step != 3 else None  # keys may come and go between steps
        theirs.update(ref_losses.TrainingLosses(backward, extra)); ours.update(my_losses.TrainingLosses(backward, extra))
    close(theirs.compute().as_dict(), ours.compute().as_dict())
    theirs.clear(); ours.clear()
    theirs.update(ref_losses.TrainingLosses(torch.tensor(2.0))); ours.update(my_losses.TrainingLosses(torch.tensor(2.0)))
    close(theirs.compute().as_dict(), ours.compute().as_dict())
    agreed += 1
    theirs = ref_losses.LossMeter(getattr(ref_losses.LossMeterType, meter_type), ref_losses.EvaluationLosses)
    ours = my_losses.LossMeter(getattr(my_losses.LossMeterType, meter_type), my_losses.EvaluationLosses)
    for step in range(4):
        checkpoint, extra = torch.rand(()), {"aux": torch.rand(())}
        theirs.update(ref_losses.EvaluationLosses(checkpoint, extra)); ours.update(my_losses.EvaluationLosses(checkpoint, extra))
    close(theirs.compute().as_dict(), ours.compute().as_dict())
    agreed += 1

# -- checkpointers: same save decisions over a trajectory, files readable by the other side -----------------------------
trajectory = [(0.9, {"val - prediction - acc": 0.50}), (0.7, {"val - prediction - acc": 0.55}), (0.8, {"val - prediction - acc": 0.70}),
              (0.6, {"val - prediction - acc": 0.65}), (0.65, {"val - prediction - acc": 0.72})]
with tempfile.TemporaryDirectory() as scratch:
    makers = [
        ("latest", lambda m, d, n: m.LatestTorchModuleCheckpointer(d, n)),
        ("best_loss", lambda m, d, n: m.BestLossTorchModuleCheckpointer(d, n)),
        ("best_metric_max", lambda m, d, n: m.BestMetricTorchModuleCheckpointer(d, n, metric="acc", maximize=True)),
        ("best_metric_min", lambda m, d, n: m.BestMetricTorchModuleCheckpointer(d, n, metric="acc", maximize=False)),
        ("function", lambda m, d, n: m.FunctionTorchModuleCheckpointer(d, n, lambda loss, metrics: loss - metrics["val - prediction - acc"], "loss minus accuracy", maximize=False)),
    ]
    for label, make in makers:
        theirs, ours = make(ref_ckpt, scratch, f"{label}_ref.pkl"), make(my_ckpt, scratch, f"{label}_mine.pkl")
        for step, (loss, metrics) in enumerate(trajectory):
            model = nn.Linear(3, 2)
            with torch.no_grad():
                model.weight.fill_(float(step))
            theirs.maybe_checkpoint(model, loss, dict(metrics)); ours.maybe_checkpoint(model, loss, dict(metrics))
            saved_ref, saved_mine = theirs.load_checkpoint(), ours.load_checkpoint()
            assert torch.equal(saved_ref.weight, saved_mine.weight), (label, step)  # the same step's model is on disk
        assert theirs.best_score == ours.best_score or abs(theirs.best_score - ours.best_score) < 1e-9
        # cross-read: each implementation loads the other's file
        assert torch.equal(theirs.load_checkpoint(str(Path(scratch) / f"{label}_mine.pkl")).weight, ours.load_checkpoint(str(Path(scratch) / f"{label}_ref.pkl")).weight)
        agreed += 1

    # -- JSON reporter: the document written for the same sequence of reports ------------------------------------------
    documents = []
    for module, folder in ((ref_report, Path(scratch) / "r"), (my_report, Path(scratch) / "m")):
        reporter = module.JsonReporter(run_id="run", output_folder=folder)
        reporter.initialize(id="client_0", name="client_0")
        reporter.report({"host_type": "client", "initialized": "2025-01-01"})
        for fl_round in (1, 2):
            reporter.report({"fit_start": f"s{fl_round}", "fit_metrics": {"acc": 0.1 * fl_round}}, round=fl_round)
            reporter.report({"fit_step_losses": 1.0}, round=fl_round, epoch=0, step=3)  # per-step data is ignored
            reporter.report({"fit_end": f"e{fl_round}", "eval_loss": 0.5 / fl_round}, round=fl_round)
            reporter.report({"fit_metrics": {"f1": 0.2}}, round=fl_round)  # nested dictionaries merge
        reporter.report({"shutdown": "now"})
        reporter.shutdown()
        files = sorted(folder.glob("*.json"))
        assert [f.name for f in files] == ["run.json"], files
        documents.append(json.loads(files[0].read_text()))
    assert documents[0] == documents[1], documents
    agreed += 1

    # -- warm start from a pretrained model, with and without a name mapping -------------------------------------------
    class Pretrained(nn.Module):
        def __init__(self) -> None:
            super().__init__()
            self.features = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 4))
            self.classifier = nn.Linear(4, 3)

    class Target(nn.Module):
        def __init__(self) -> None:
            super().__init__()
            self.base_module = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 4))
            self.head = nn.Linear(4, 2)  # different width: must be left alone
            self.features = nn.Sequential(nn.Linear(6, 5))

    pretrained = Pretrained()
    mapping_path = Path(scratch) / "mapping.json"
    mapping_path.write_text(json.dumps({"base_module": "features", "head": "classifier"}))
    torch.save(pretrained, Path(scratch) / "pretrained.pt")
    for kwargs in ({"pretrained_model": pretrained}, {"pretrained_model": pretrained, "weights_mapping_path": mapping_path},
                   {"pretrained_model_path": Path(scratch) / "pretrained.pt", "weights_mapping_path": mapping_path}):
        torch.manual_seed(9); target_ref = Target()
        torch.manual_seed(9); target_mine = Target()
        loaded_ref = ref_warm.WarmedUpModule(**kwargs).load_from_pretrained(target_ref)
        loaded_mine = my_warm.WarmedUpModule(**kwargs).load_from_pretrained(target_mine)
        for (name, a), (_, b) in zip(loaded_ref.state_dict().items(), loaded_mine.state_dict().items()):
            assert torch.equal(a, b), name
        probe_ref, probe_mine = ref_warm.WarmedUpModule(**kwargs), my_warm.WarmedUpModule(**kwargs)
        for key in ("base_module.0.weight", "head.bias", "features.0.weight", "unknown.weight"):
            assert probe_ref.get_matching_component(key) == probe_mine.get_matching_component(key), key
        agreed += 1
print("configs agree:", agreed)
