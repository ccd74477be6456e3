"""The multi-process runtime vs the reference: the same scenarios as ``check_federations.py``, but this framework runs them
SPMD -- one process per client (``torch.distributed.run``, gloo on CPU), server logic replicated on every rank, payloads
reduced with collectives -- while the reference runs its server / clients over its transport.  Histories must coincide."""
import json
import os
import socket
import subprocess
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

sys.path.insert(0, str(Path(__file__).resolve().parent))
import check_federations  # noqa: E402
from check_federations import CLIENTS, ROUNDS, SCENARIOS, build, compare, resolver, run_reference  # noqa: E402

# all of these agree (python check_spmd_federations.py all); the default run keeps the suite short and covers the distinct
# payload kinds: plain weights, weights ++ scalar, weights ++ variates (+ warm start), partial exchange, per-client
# aggregation weights, server-side optimizer state
ALL_SPMD_SCENARIOS = ["fedavg", "fedprox", "ditto", "scaffold", "apfl", "moon", "fedper", "fedbn", "fenda", "gpfl", "feddg_ga", "flash"]
SPMD_SCENARIOS = ["fedprox", "scaffold", "fedper", "feddg_ga"]


def worker(name: str, out_path: str) -> None:
    from fl4health_b200.parallel.spmd import SpmdContext, build_spmd_federation
    from fl4health_b200.parallel.spmd_multi import build_spmd_federation_multi

    ctx = SpmdContext()
    server, clients = build(resolver("fl4health_b200"), SCENARIOS[name], ours=True)
    if ctx.world_size == CLIENTS:
        build_spmd_federation(ctx, server, clients[ctx.rank], fused=False)
    else:  # fewer ranks than clients: uneven hosting (rank 0 takes the remainder), two-level reduction
        per_rank = CLIENTS // ctx.world_size
        first = ctx.rank * per_rank + (CLIENTS % ctx.world_size if ctx.rank else 0)
        count = per_rank + (CLIENTS % ctx.world_size if ctx.rank == 0 else 0)
        build_spmd_federation_multi(ctx, server, clients[first:first + count])
    history, _ = server.fit(num_rounds=ROUNDS)
    if ctx.rank == 0:
        Path(out_path).write_text(json.dumps({
            "losses_distributed": history.losses_distributed, "metrics_distributed_fit": history.metrics_distributed_fit,
            "metrics_distributed": history.metrics_distributed}))
    ctx.barrier()
    ctx.shutdown()


def run_spmd(name: str, ranks: int = CLIENTS) -> SimpleNamespace:
    with socket.socket() as probe:
        probe.bind(("127.0.0.1", 0))
        port = probe.getsockname()[1]
    out = Path(tempfile.mkdtemp(prefix="spmd_")) / f"{name}.json"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", FL4H_LOG_LEVEL="ERROR")
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), str(Path(__file__).resolve()), "--worker", name, str(out)]
    done = subprocess.run(command, env=env, capture_output=True, text=True, timeout=600)
    assert done.returncode == 0, (name, done.stdout[-2000:], done.stderr[-3000:])
    data = json.loads(out.read_text())
    as_pairs = lambda series: [(int(r), v) for r, v in series]  # noqa: E731
    return SimpleNamespace(losses_distributed=as_pairs(data["losses_distributed"]),
                           metrics_distributed_fit={k: as_pairs(v) for k, v in data["metrics_distributed_fit"].items()},
                           metrics_distributed={k: as_pairs(v) for k, v in data["metrics_distributed"].items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(sys.argv[2], sys.argv[3])
    else:
        wanted = ALL_SPMD_SCENARIOS if sys.argv[1:] == ["all"] else (sys.argv[1:] or SPMD_SCENARIOS)
        for name in wanted:
            compare(f"spmd:{name}", run_reference(SCENARIOS[name]), run_spmd(name), tol=2e-4)
        if not sys.argv[1:] or sys.argv[1:] == ["all"]:
            for name in (("fedavg", "scaffold") if sys.argv[1:] == ["all"] else ("fedavg",)):  # three clients on two processes ([2, 1])
                compare(f"spmd, 2 ranks host 3 clients:{name}", run_reference(SCENARIOS[name]), run_spmd(name, ranks=2), tol=2e-4)
        print("configs agree:", check_federations.agreed)
