#!/usr/bin/env python
"""Reference arm of ``bench.py``: the UNMODIFIED FL4Health package (``baseline/_ref/fl4health``, a byte-for-byte copy of
``/root/reference/fl4health``) running its stock FedAvg path for the headline config.

Nothing in this file (or in ``baseline/stubs``) imports ``fl4health_b200``.  The deployment shape is the reference's
own: ONE server process (``FlServer`` + ``BasicFedAvg`` on the CPU, NumPy aggregation) and ONE client process per GPU
(``BasicClient`` + ``FullParameterExchanger``), talking over localhost TCP through the stand-alone ``flwr`` shim (the
image has no Flower wheel; see ``baseline/stubs/flwr/__init__.py``).  Every round therefore does what the reference
does: ``state_dict -> .cpu().numpy() -> np.save bytes -> wire -> np.load -> aggregate() in NumPy -> np.save -> wire ->
np.load -> torch.tensor -> load_state_dict`` (``fl4health/parameter_exchange/full_exchanger.py:30,45-47``,
``fl4health/strategies/aggregate_utils.py:8-55``), eager fp32 PyTorch local training with a stock
``torch.utils.data.DataLoader`` (``fl4health/clients/basic_client.py:294-386``), PyTorch's default cuDNN-TF32 conv
math, no AMP.

Workload = ``bench.py``'s: ResNet-18 (CIFAR stem, 11.17 M params, torchvision ``ResNet(BasicBlock,[2,2,2,2])`` with a
3x3 stem and no max-pool), synthetic CIFAR-10-shaped data, ``local_steps`` SGD(lr 0.01, momentum 0.9) steps of batch
32, then ``val_batches`` validation batches, per client per round.  The session runs ``2 x (warmup + steps)`` rounds:
the first half with the datasets on the host (e2e: every batch is an H2D copy from pinned DataLoader memory, every
round's loss/metrics are read back), the second half with the dataset tensors already resident on the GPU (``value``).
Timed with CUDA events on every client rank between a barrier + synchronize on both sides, max over ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
for extra in (HERE / "stubs", HERE / "_ref"):
    if str(extra) not in sys.path:
        sys.path.insert(0, str(extra))
os.environ.setdefault("FLWR_SHIM_LOG_LEVEL", "WARNING")


def reference_available() -> str | None:
    """None when the reference package is importable, else a one-line reason."""
    target = HERE / "_ref" / "fl4health"
    source = Path(os.environ.get("FL4H_REFERENCE_SRC", "/root/reference")) / "fl4health"
    if not (target / "__init__.py").exists() and (source / "__init__.py").exists():
        import shutil  # same outcome as baseline/install_reference.sh: an unmodified copy of the pure-Python package

        shutil.copytree(source, target, ignore=shutil.ignore_patterns("__pycache__"))
    if not (target / "__init__.py").exists():
        return ("baseline/_ref/fl4health is missing: `pip install --target baseline/_ref /root/reference` fails offline "
                "(no hatchling, requires-python <3.11) and the fallback copy (baseline/install_reference.sh) was not run")
    return None


def parse_args(argv: list[str] | None = None) -> argparse.Namespace:
    p = argparse.ArgumentParser()
    p.add_argument("--role", default="client", choices=["client", "server"])
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--local-steps", type=int, default=8)
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--val-batches", type=int, default=4)
    p.add_argument("--train-samples", type=int, default=4096)
    p.add_argument("--port", type=int, default=0)
    p.add_argument("--device", default="cuda", help="cuda | cpu (cpu is for the plumbing test only)")
    p.add_argument("--skip-e2e", action="store_true")
    return p.parse_known_args(argv)[0]


def session_rounds(args: argparse.Namespace) -> int:
    return (args.warmup + args.steps) * (1 if args.skip_e2e else 2)


# ----------------------------------------------------------------------------------------------------- server process
def run_server(args: argparse.Namespace) -> None:
    import flwr
    from fl4health.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn
    from fl4health.servers.base_server import FlServer
    from fl4health.strategies.basic_fedavg import BasicFedAvg
    from flwr.server.client_manager import SimpleClientManager

    def config_fn(server_round: int) -> dict:
        return {"current_server_round": server_round, "local_steps": args.local_steps, "batch_size": args.batch_size}

    n = args.gpus
    strategy = BasicFedAvg(
        min_fit_clients=n, min_evaluate_clients=n, min_available_clients=n,
        on_fit_config_fn=config_fn, on_evaluate_config_fn=config_fn,
        fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
        evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn,
    )
    server = FlServer(
        client_manager=SimpleClientManager(), fl_config={"n_server_rounds": session_rounds(args)}, strategy=strategy,
        on_init_parameters_config_fn=config_fn, accept_failures=False,
    )
    history = flwr.server.start_server(
        server=server, server_address=f"127.0.0.1:{args.port}",
        config=flwr.server.ServerConfig(num_rounds=session_rounds(args)),
    )
    server.shutdown()
    final = history.losses_distributed[-1][1] if history.losses_distributed else None
    print(json.dumps({"server_final_val_loss": final}), file=sys.stderr)


# ----------------------------------------------------------------------------------------------------- clock sampling
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.proc = None
        self.gpu_index = gpu_index

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------- client process
def build_resnet18_cifar():  # noqa: ANN201
    """torchvision's ResNet-18 with the CIFAR stem (3x3 conv, no max-pool), 10 classes: 11,173,962 parameters."""
    from torch import nn
    from torchvision.models.resnet import BasicBlock, ResNet

    model = ResNet(BasicBlock, [2, 2, 2, 2], num_classes=10)
    model.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
    nn.init.kaiming_normal_(model.conv1.weight, mode="fan_out", nonlinearity="relu")
    model.maxpool = nn.Identity()
    return model


def run_client(args: argparse.Namespace) -> None:
    import torch
    import torch.distributed as dist
    from torch import nn
    from torch.utils.data import DataLoader

    import flwr
    from fl4health.clients.basic_client import BasicClient
    from fl4health.metrics import Accuracy
    from fl4health.utils.dataset import TensorDataset

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    use_cuda = args.device == "cuda"
    if use_cuda:
        assert torch.cuda.is_available(), "the reference arm needs a CUDA device (use --device cpu for a plumbing test)"
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
    else:
        device = torch.device("cpu")
    if world > 1:
        dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=world,
                                **({"device_id": device} if use_cuda else {}))

    # rank 0 owns the server process; the port is agreed through the environment (MASTER_PORT + 1017) or picked free
    if args.port == 0:
        if world > 1:
            args.port = int(os.environ.get("MASTER_PORT", "29500")) + 1017
        else:
            with socket.socket() as s:
                s.bind(("127.0.0.1", 0))
                args.port = s.getsockname()[1]
    server_proc = None
    if rank == 0:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")  # the reference server aggregates on the CPU
        cmd = [sys.executable, str(Path(__file__).resolve()), "--role", "server", "--gpus", str(args.gpus), "--steps",
               str(args.steps), "--warmup", str(args.warmup), "--local-steps", str(args.local_steps), "--batch-size",
               str(args.batch_size), "--port", str(args.port)] + (["--skip-e2e"] if args.skip_e2e else [])
        server_proc = subprocess.Popen(cmd, env=env)

    torch.backends.cudnn.benchmark = True  # what the reference recommends for its own GPU runs (nnunet_client.py:203)
    torch.manual_seed(1234 + rank)

    def synthetic(n: int, seed: int) -> tuple[torch.Tensor, torch.Tensor]:
        gen = torch.Generator().manual_seed(seed)
        targets = torch.randint(0, 10, (n,), generator=gen)
        data = torch.randn(n, 3, 32, 32, generator=gen) * 0.5 + (targets.float().view(-1, 1, 1, 1) - 4.5) * 0.1
        return data, targets

    first_phase = args.warmup + args.steps  # rounds [1, first_phase] = e2e, (first_phase, 2*first_phase] = resident
    marks: dict[str, object] = {}
    results: dict[str, dict] = {}

    def barrier() -> None:
        if world > 1:
            dist.barrier()

    def sync() -> None:
        if use_cuda:
            torch.cuda.synchronize()

    class Stamp:
        """CUDA event on GPU runs, perf_counter on the CPU plumbing test."""

        def __init__(self) -> None:
            self.wall = time.perf_counter()
            self.event = None
            if use_cuda:
                self.event = torch.cuda.Event(enable_timing=True)
                self.event.record()

        def ms_until(self, other: "Stamp") -> float:
            if self.event is not None:
                return self.event.elapsed_time(other.event)
            return (other.wall - self.wall) * 1e3

    class ReferenceClient(BasicClient):
        resident = args.skip_e2e

        def get_model(self, config):  # noqa: ANN001, ANN202
            torch.manual_seed(1234)
            return build_resnet18_cifar().to(self.device)

        def make_loaders(self, batch_size: int):  # noqa: ANN202
            train_x, train_y = synthetic(args.train_samples, 100 + rank)
            val_x, val_y = synthetic(args.val_batches * batch_size, 900 + rank)
            if self.resident:
                train_x, train_y, val_x, val_y = (t.to(self.device) for t in (train_x, train_y, val_x, val_y))
            pin = use_cuda and not self.resident
            train = DataLoader(TensorDataset(train_x, train_y), batch_size=batch_size, shuffle=True, drop_last=True,
                               pin_memory=pin)
            val = DataLoader(TensorDataset(val_x, val_y), batch_size=batch_size, pin_memory=pin)
            return train, val

        def get_data_loaders(self, config):  # noqa: ANN001, ANN202
            return self.make_loaders(int(config["batch_size"]))

        def get_criterion(self, config):  # noqa: ANN001, ANN202
            return nn.CrossEntropyLoss()

        def get_optimizer(self, config):  # noqa: ANN001, ANN202
            return torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9)

        # --- harness only: phase switch + timing marks around the reference's own fit()/evaluate() ---------------
        def fit(self, parameters, config):  # noqa: ANN001, ANN202
            server_round = int(config["current_server_round"])
            if not args.skip_e2e and server_round == first_phase + 1:
                type(self).resident = True
                self.train_loader, self.val_loader = self.make_loaders(int(config["batch_size"]))
                self.train_iterator = iter(self.train_loader)
            phase_round = (server_round - 1) % first_phase + 1
            if phase_round == args.warmup + 1:
                barrier()
                sync()
                if rank == 0 and use_cuda:
                    marks["sampler"] = ClockSampler(local_rank)
                    marks["sampler"].start()
                marks["start"] = Stamp()
            return super().fit(parameters, config)

        def evaluate(self, parameters, config):  # noqa: ANN001, ANN202
            out = super().evaluate(parameters, config)
            server_round = int(config["current_server_round"])
            phase_round = (server_round - 1) % first_phase + 1
            if phase_round == args.warmup + args.steps:
                end = Stamp()
                sync()
                barrier()
                label = "resident" if type(self).resident else "e2e"
                ms = marks["start"].ms_until(end)
                results[label] = {"ms_total": ms, "final_loss": float(out[0])}
                sampler = marks.pop("sampler", None)
                if sampler is not None:
                    results[label]["clocks"] = sampler.stop()
            return out

    client = ReferenceClient(Path("."), [Accuracy()], device, client_name=f"ref_rank{rank}")
    flwr.client.start_client(server_address=f"127.0.0.1:{args.port}", client=client.to_client(), cid=f"rank{rank}")
    client.shutdown()
    if server_proc is not None:
        server_proc.wait(timeout=120)

    def max_over_ranks(value: float) -> float:
        if world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    main = results["resident"]
    ms_per_round = max_over_ranks(main["ms_total"]) / args.steps
    e2e_ms = None if args.skip_e2e else max_over_ranks(results["e2e"]["ms_total"]) / args.steps
    per_round_samples = (args.local_steps + args.val_batches) * args.batch_size
    bytes_in = per_round_samples * (3 * 32 * 32 * 4 + 8)
    payload = sum(v.numel() * v.element_size() for v in client.model.state_dict().values())
    line = {
        "impl": "reference",
        "metric": "fl_rounds_per_sec_cifar10_resnet18_fedavg",
        "value": world * 1000.0 / ms_per_round,
        "unit": "client-rounds/s (= FL rounds/s x N clients; FL rounds/s at N=1)",
        "federation_rounds_per_s": 1000.0 / ms_per_round,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_round,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32 (PyTorch defaults: cuDNN conv TF32 allowed, fp32 matmul)",
        "data": "synthetic CIFAR-10-shaped (3x32x32, 10 classes), random-init ResNet-18",
        "config": {
            "model": "resnet18_cifar (11.17M params)", "clients": world,
            "parallelism": f"fl_dp{world} (one client process per GPU + one CPU server process)",
            "global_batch": args.batch_size * world, "batch_per_client": args.batch_size,
            "local_steps": args.local_steps, "val_batches_per_client": args.val_batches,
            "strategy": "BasicFedAvg (weighted)", "optimizer": "SGD lr=0.01 momentum=0.9", "seq_len": None,
            "l2": "inputs > L2 per round: every round moves the 44.7 MB model through host memory twice",
            "collectives": "reference stock path: state_dict -> numpy -> np.save bytes -> localhost TCP -> NumPy aggregate",
            "reference_package": "baseline/_ref/fl4health (unmodified copy of /root/reference/fl4health)",
            "transport": "baseline/stubs/flwr shim (no Flower wheel in the image)",
            "exchange_payload_bytes": payload,
        },
        "gpu_launches": 0,
        "clocks": main.get("clocks"),
        "final_val_loss": main["final_loss"],
    }
    if e2e_ms is not None:
        line["e2e"] = {
            "value": world * 1000.0 / e2e_ms, "unit": "client-rounds/s (same aggregate as `value`)",
            "federation_rounds_per_s": 1000.0 / e2e_ms, "ms_per_step": e2e_ms,
            "h2d_bytes_per_step": (bytes_in + payload) * world, "d2h_bytes_per_step": (payload + 6 * 4) * world,
            "note": "stock DataLoader over host tensors (pin_memory), H2D per batch; parameters cross PCIe both ways "
                    "every round (reference design)",
        }
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main(argv: list[str] | None = None) -> None:
    args = parse_args(argv)
    reason = reference_available()
    if reason is not None:
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": reason}))
        return
    if args.role == "server":
        run_server(args)
    else:
        run_client(args)


if __name__ == "__main__":
    main()
