"""What the reporters record, vs the reference: the same federation with a ``JsonReporter`` on the server and on every
client.  The documents must have the same structure (every key the reference writes, at the same nesting) and the same
values wherever the value is not a wall-clock reading."""
import json
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import check_federations  # noqa: E402
from check_federations import SCENARIOS, resolver, run_ours, run_reference  # noqa: E402

CLOCK_KEYS = ("start", "end", "time", "elapsed", "initialized", "shutdown")  # wall-clock readings: present, not compared


def with_reporters(scenario: dict, folder: Path) -> dict:
    scenario = dict(scenario)
    inner_client_args = scenario.get("client_args", lambda side: {})
    inner_server_args = scenario.get("server_args", lambda side: {})
    counter = {"n": 0}

    def client_args(side):
        counter["n"] += 1
        return {**inner_client_args(side), "reporters": [side("reporting.json_reporter").JsonReporter(run_id=f"client_{counter['n'] - 1}", output_folder=folder)]}

    def server_args(side):
        return {**inner_server_args(side), "reporters": [side("reporting.json_reporter").JsonReporter(run_id="server", output_folder=folder)]}

    scenario["client_args"], scenario["server_args"] = client_args, server_args
    return scenario


def is_clock(key: str) -> bool:
    return any(word in key for word in CLOCK_KEYS)


def same_document(path: str, theirs, ours) -> int:
    """Every key of the reference's document exists in ours with an equal value (clock readings: same type)."""
    checked = 0
    if isinstance(theirs, dict):
        assert isinstance(ours, dict), (path, type(ours))
        missing = [k for k in theirs if k not in ours]
        assert not missing, (path, "keys missing from our report", missing, sorted(ours))
        for key, value in theirs.items():
            checked += same_document(f"{path}/{key}", value, ours[key])
        return checked
    if is_clock(path.rsplit("/", 1)[-1]):
        assert type(theirs) is type(ours) or isinstance(ours, (int, float, str)), (path, theirs, ours)
        return 1
    if isinstance(theirs, float) or isinstance(ours, float):
        assert abs(float(theirs) - float(ours)) <= 2e-4 * max(1.0, abs(float(theirs))), (path, theirs, ours)
    else:
        assert theirs == ours, (path, theirs, ours)
    return 1


if __name__ == "__main__":
    total = 0
    for name in ("fedavg", "fedprox", "ditto", "fedavg_unweighted_eval_after_fit"):
        documents = {}
        for label, run in (("reference", run_reference), ("ours", run_ours)):
            folder = Path(tempfile.mkdtemp(prefix=f"reports_{label}_"))
            run(with_reporters(SCENARIOS[name], folder))
            documents[label] = {path.name: json.loads(path.read_text()) for path in sorted(folder.glob("*.json"))}
        assert documents["reference"].keys() == documents["ours"].keys(), (name, sorted(documents["reference"]), sorted(documents["ours"]))
        for file_name, theirs in documents["reference"].items():
            total += same_document(f"{name}/{file_name}", theirs, documents["ours"][file_name])
        check_federations.agreed += 1
    print(f"  {total} report entries compared", file=sys.stderr)
    print("configs agree:", check_federations.agreed)
