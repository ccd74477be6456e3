#!/usr/bin/env python
"""Headline benchmark: FL rounds/sec, CIFAR-10-shaped ResNet-18 FedAvg, one client per GPU (BASELINE.json).

``value`` is the whole-job aggregate the driver contract asks for: client rounds per second summed over the N
client-GPUs (= N x the federation's rounds/s; the two coincide at N=1).  The federation's own rounds/s is reported next
to it as ``federation_rounds_per_s`` and ``ms_per_step`` is the duration of one federation round.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 --steps 20 --warmup 3

One "step" = one full federated round driven through the public API (``FlServer.fit``): server->client parameter
exchange, ``local_steps`` SGD(momentum) steps of batch 32 on each client GPU, client->server weighted FedAvg
aggregation, then a federated evaluation pass (``val_batches`` batches per client) with metric aggregation.
Synthetic CIFAR-10-shaped data, random-init ResNet-18 (there is no network for datasets).  Timed with CUDA events
on every rank (barrier + synchronize on both sides), max over ranks.  Two timed regions:

* ``value``  — datasets resident in HBM (no host traffic except the per-round scalar read-back);
* ``e2e``    — every batch is staged from pinned host memory (H2D inside the timed region) and every round's loss /
  metrics are read back (D2H), through the same public API.

Precision: the headline (``value`` / ``e2e``) runs at the REFERENCE's precision — fp32 parameters, activations and
optimizer state, PyTorch-default TF32 tensor-core math inside convolutions, fp32 Linear (the reference has no AMP
outside nnU-Net).  The bf16 master-weight engine mode is measured afterwards in the same process and reported under
``"bf16"`` as an extra (``--dtype bf16`` makes it the headline instead).

``--impl reference`` runs the UNMODIFIED reference package (``baseline/_ref/fl4health``) through its stock path
(``baseline/reference_arm.py``: one CPU server process + one client process per GPU, NumPy aggregation, no
``fl4health_b200`` import anywhere on that path); see DESIGN.md §6.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")
sys.path.insert(0, str(Path(__file__).resolve().parent))


def parse_args() -> argparse.Namespace:
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200, help="timed FL rounds")
    p.add_argument("--warmup", type=int, default=5, help="untimed warm-up FL rounds")
    p.add_argument("--impl", default="native", choices=["native", "reference", "eager"])
    p.add_argument("--local-steps", type=int, default=8)
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--val-batches", type=int, default=4)
    p.add_argument("--train-samples", type=int, default=4096)
    p.add_argument("--collectives", default="auto", choices=["auto", "nccl", "fused"])
    p.add_argument("--no-graphs", action="store_true")
    p.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"],
                   help="headline precision: fp32 = the reference's (fp32 state, TF32 conv math); bf16 = master-weight mode")
    p.add_argument("--fp32", action="store_true", help="(kept for old command lines) same as --dtype fp32")
    p.add_argument("--skip-extra-dtype", action="store_true", help="do not also measure the other precision")
    p.add_argument("--no-master-weights", action="store_true", help="bf16 autocast over fp32 params instead of bf16 shadow params")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--device", default="cuda", help="reference arm only: cpu runs its plumbing test without a GPU")
    p.add_argument("--config", default="cifar_fedavg", choices=["cifar_fedavg", "scaffold_fedprox", "fedper_ditto_dp", "bert_fedadam"],
                   help="BASELINE.json configuration (default: the headline CIFAR-10 ResNet-18 FedAvg)")
    return p.parse_args()


def reference_arm() -> None:
    """Hand over to baseline/reference_arm.py (imports nothing from fl4health_b200; same CLI flags)."""
    import runpy

    sys.path.pop(0)  # the reference arm must not see this repo's package
    runpy.run_path(str(Path(__file__).resolve().parent / "baseline" / "reference_arm.py"), run_name="__main__")


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.proc = None
        self.gpu_index = gpu_index

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
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
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(sm_max) if sm_max else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        if args.config != "cifar_fedavg":
            print(json.dumps({"impl": "reference", "config": args.config,
                              "unavailable": "the reference arm drives the headline configuration (cifar_fedavg) only"}))
            return
        reference_arm()
        return
    if args.config == "bert_fedadam":  # BASELINE config #4 has its own harness (token batches, FedAdam server, bf16)
        import runpy

        sys.argv = [sys.argv[0], "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        runpy.run_path(str(Path(__file__).resolve().parent / "benchmarks" / "bert_fedopt_round.py"), run_name="__main__")
        return

    import torch
    from torch import nn

    from fl4health_b200 import ops
    from fl4health_b200.engine.data import BatchedTensorLoader
    from fl4health_b200.engine.options import EngineOptions
    from fl4health_b200.parallel.spmd import SpmdContext, build_spmd_federation
    from fl4health_b200.utils.dataset import TensorDataset
    from fl4health_b200.utils import tracing

    sys.path.insert(0, str(Path(__file__).resolve().parent / "benchmarks"))
    import fl_variants

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    os.environ["FL4H_COLLECTIVES"] = args.collectives
    ctx = SpmdContext()
    world = ctx.world_size
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    device = ctx.device
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234 + ctx.rank)

    eager_impl = args.impl == "eager"
    headline_dtype = "fp32" if args.fp32 else args.dtype

    def engine_for(dtype: str) -> EngineOptions:
        bf16 = dtype == "bf16"
        return EngineOptions(
            arena=not eager_impl, fused_optimizer=not eager_impl, cuda_graphs=not (args.no_graphs or eager_impl),
            amp_dtype=torch.bfloat16 if bf16 else None, channels_last=not eager_impl,
            master_weights=bf16 and not (eager_impl or args.no_master_weights),
            table_grads=not (bf16 or eager_impl) and os.environ.get("FL4H_TABLE_GRADS", "1") != "0",  # fp32: pointer-table optimizer
        )

    def synthetic(n: int, seed: int) -> TensorDataset:
        gen = torch.Generator().manual_seed(seed)
        targets = torch.randint(0, 10, (n,), generator=gen)
        data = torch.randn(n, 3, 32, 32, generator=gen) * 0.5 + (targets.float().view(-1, 1, 1, 1) - 4.5) * 0.1
        return TensorDataset(data, targets)

    class BenchHooks:
        """Data / criterion / optimizer factories shared by every variant's client class (mixed in before the
        algorithm's client, see benchmarks/fl_variants.py)."""

        placement = "device"

        def get_data_loaders(self, config):  # noqa: ANN001, ANN202
            bs = int(config["batch_size"])
            train = BatchedTensorLoader(synthetic(args.train_samples, 100 + ctx.rank), bs, shuffle=True, drop_last=True,
                                        placement=BenchHooks.placement, device=self.device)
            val = BatchedTensorLoader(synthetic(args.val_batches * bs, 900 + ctx.rank), bs, placement=BenchHooks.placement,
                                      device=self.device)
            return train, val

        def get_criterion(self, config):  # noqa: ANN001, ANN202
            return nn.CrossEntropyLoss()

        def get_optimizer(self, config):  # noqa: ANN001, ANN202
            return torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9)

    def config_fn(server_round: int) -> dict:
        return {"current_server_round": server_round, "local_steps": args.local_steps, "batch_size": args.batch_size}

    l2_flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=device)  # 256 MiB > 126 MB L2

    def run_federation(dtype: str, with_e2e: bool, variant: str = "fedavg") -> tuple[dict, dict | None, EngineOptions]:
        """Build one federation at `dtype` and time it (e2e staging first, then device-resident datasets)."""
        engine = engine_for(dtype)

        def federation(placement: str):  # noqa: ANN202
            BenchHooks.placement = placement
            client, server = fl_variants.build(variant, BenchHooks, ctx, engine, args.steps + args.warmup, args.local_steps, args.batch_size)
            build_spmd_federation(ctx, server, client, fused=False if args.collectives == "nccl" else None)
            return client, server

        client, server = federation("pinned" if with_e2e else "device")

        def timed_fit(label: str) -> dict:
            """Run warmup+steps rounds through FlServer.fit; time the last `steps` rounds on the device."""
            marks: dict[str, object] = {}
            sampler = ClockSampler(device.index) if ctx.rank == 0 else None

            def on_round_end(server_round: int) -> None:
                l2_flush.fill_(1.0)  # evict L2 between rounds (inside the timed region; ~40 us)
                if server_round == args.warmup:
                    ctx.barrier()
                    torch.cuda.synchronize()
                    if sampler is not None:
                        sampler.start()
                    ops.reset_launch_count()
                    tracing.phase_report(reset=True)  # FL4H_TRACE=1 diagnostics cover the timed rounds only
                    marks["start"] = torch.cuda.Event(enable_timing=True)
                    marks["start"].record()
                    marks["wall0"] = time.perf_counter()
                elif server_round == args.warmup + args.steps:
                    marks["end"] = torch.cuda.Event(enable_timing=True)
                    marks["end"].record()
                    torch.cuda.synchronize()
                    ctx.barrier()
                    marks["wall1"] = time.perf_counter()
                    marks["launches"] = ops.launch_count()
                    if sampler is not None:
                        marks["clocks"] = sampler.stop()

            server.round_end_hooks = [on_round_end]
            if args.warmup == 0:
                on_round_end(0)
            history, _ = server.fit(num_rounds=args.warmup + args.steps)
            ms = marks["start"].elapsed_time(marks["end"])  # type: ignore[union-attr]
            ms = ctx.all_reduce_max(ms)
            final_loss = history.losses_distributed[-1][1]
            return {"label": label, "ms_total": ms, "ms_per_round": ms / args.steps, "launches": marks["launches"],
                    "clocks": marks.get("clocks"), "final_loss": final_loss,
                    "wall_s": marks["wall1"] - marks["wall0"]}  # type: ignore[operator]

        # ---- e2e first (pinned host datasets -> H2D every batch), then device-resident ------------------------
        e2e_run = None
        if with_e2e:
            e2e_run = timed_fit("e2e")
            if variant == "fedavg":  # re-create loaders for the device-resident run (graphs and model state are reused)
                BenchHooks.placement = "device"
                client.train_loader, client.val_loader = client.get_data_loaders(config_fn(1))
                client.train_iterator = None
            else:
                # algorithms with server-held initial state (SCAFFOLD's variates) or partial exchange (FedPer) define
                # one fit() per federation: the device-resident pass gets a fresh client and server
                client, server = federation("device")
        device_run = timed_fit("device")
        if tracing.tracing_enabled() and ctx.rank == 0:  # FL4H_TRACE=1: device ms per round phase (diagnostic, stderr)
            torch.cuda.synchronize()
            report = tracing.phase_report()
            print(json.dumps({"dtype": dtype, **{k: round(v["mean_ms"], 4) for k, v in report.items()}}), file=sys.stderr)
        return device_run, e2e_run, engine

    variants = fl_variants.GROUPS[args.config]
    main_run, e2e, engine = run_federation(headline_dtype, with_e2e=not args.skip_e2e, variant=variants[0])
    others = {}
    for name in variants[1:]:  # further algorithms of this BASELINE configuration: device-resident timing each
        run, _, _ = run_federation(headline_dtype, with_e2e=False, variant=name)
        others[name] = {"value": world * 1000.0 / run["ms_per_round"], "ms_per_step": run["ms_per_round"],
                        "final_val_loss": run["final_loss"], "gpu_launches": run["launches"]}
    extra = None
    if not (args.skip_extra_dtype or eager_impl or args.config != "cifar_fedavg"):
        other = "bf16" if headline_dtype == "fp32" else "fp32"
        extra_run, extra_e2e, _ = run_federation(other, with_e2e=not args.skip_e2e)
        extra = {"dtype": other, "value": world * 1000.0 / extra_run["ms_per_round"], "ms_per_step": extra_run["ms_per_round"],
                 "final_val_loss": extra_run["final_loss"], "gpu_launches": extra_run["launches"]}
        if extra_e2e is not None:
            extra["e2e"] = {"value": world * 1000.0 / extra_e2e["ms_per_round"], "ms_per_step": extra_e2e["ms_per_round"]}

    rounds_per_s = 1000.0 / main_run["ms_per_round"]
    bytes_in = args.local_steps * args.batch_size * (3 * 32 * 32 * 4 + 8) + args.val_batches * args.batch_size * (3 * 32 * 32 * 4 + 8)
    result = {
        "metric": ("fl_rounds_per_sec_cifar10_resnet18_fedavg" if args.config == "cifar_fedavg"
                   else f"fl_rounds_per_sec_cifar10_{args.config}[{variants[0]}]"),
        # whole-job aggregate: client rounds completed per second summed over the N client-GPUs.  The federation as a
        # whole advances `federation_rounds_per_s` rounds per second, each round doing N clients' worth of work (weak
        # scaling: per-GPU work fixed), so the aggregate is N x that; at N=1 both are the reference's "FL rounds/sec".
        "value": rounds_per_s * world,
        "unit": "client-rounds/s (= FL rounds/s x N clients; FL rounds/s at N=1)",
        "federation_rounds_per_s": rounds_per_s,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": main_run["ms_per_round"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": ("fp32 (fp32 params/activations/optimizer, TF32 tensor-core conv math = the reference's PyTorch defaults)"
                  if headline_dtype == "fp32" else "bf16"),
        "data": "synthetic CIFAR-10-shaped (3x32x32, 10 classes), random-init ResNet-18",
        "impl": args.impl,
        "config": {
            "model": "resnet18_cifar (11.17M params)", "clients": world, "parallelism": f"fl_dp{world} (one client per GPU)",
            "global_batch": args.batch_size * world, "batch_per_client": args.batch_size,
            "local_steps": args.local_steps, "val_batches_per_client": args.val_batches, "strategy": "BasicFedAvg (weighted)" if args.config == "cifar_fedavg" else variants[0],
            "optimizer": "SGD lr=0.01 momentum=0.9", "seq_len": None,
            "l2": "explicit 256 MiB write between rounds (inside the timed region)",
            "collectives": (("fused-nvls (multimem)" if ctx.fused.has_multicast else "fused-p2p") if ctx.fused is not None
                            else ("nccl" if world > 1 else "local")),
            "cuda_graphs": engine.cuda_graphs, "channels_last": engine.channels_last,
        },
        "gpu_launches": main_run["launches"],  # this rank's launches of fl4h kernels in the timed region (each rank: same)
        "clocks": main_run["clocks"],
        "final_val_loss": main_run["final_loss"],
        "wall_s": main_run["wall_s"],
    }
    if e2e is not None:
        result["e2e"] = {
            "value": world * 1000.0 / e2e["ms_per_round"], "unit": "client-rounds/s (same aggregate as `value`)",
            "federation_rounds_per_s": 1000.0 / e2e["ms_per_round"], "ms_per_step": e2e["ms_per_round"],
            "h2d_bytes_per_step": bytes_in * world, "d2h_bytes_per_step": 6 * 4 * world,
            "note": "per federation round, summed over ranks: every train/val batch copied from pinned host memory; "
                    "loss+accuracy scalars read back on every rank",
        }
    if others:
        result["variants"] = {variants[0]: {"value": result["value"], "ms_per_step": result["ms_per_step"]}, **others}
    if extra is not None:
        result[extra["dtype"]] = extra  # the other precision, same federation code, measured after the headline
    if ctx.rank == 0:
        print(json.dumps(result))
    ctx.shutdown()


if __name__ == "__main__":
    main()
