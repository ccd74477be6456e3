"""Accuracy parity of the execution modes on one learnable synthetic task (CIFAR-shaped images whose class shifts the
mean): the same federation -- same seeds, data, initial weights, rounds -- is trained as

    fp32-eager   every engine feature off (stock convolutions / BatchNorm / optimizer, no graphs): the reference's style
    fp32         headline configuration (own tcgen05 TF32 convolutions, fused BN, table optimizer, CUDA graphs)
    bf16         bf16 activations, fp32 master weights

and, under ``torchrun`` with several ranks, with fused peer-memory collectives vs NCCL (``--collectives``).  Prints one
``PARITY {json}`` line per mode with the per-round federated validation loss / accuracy.

    python benchmarks/convergence_parity.py --rounds 30
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/convergence_parity.py --modes fp32 --collectives auto nccl
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "ERROR")

import torch  # noqa: E402
from torch import nn  # noqa: E402

import fl_variants  # noqa: E402

from fl4health_b200.engine.data import BatchedTensorLoader  # noqa: E402
from fl4health_b200.engine.options import EngineOptions  # noqa: E402
from fl4health_b200.parallel.spmd import SpmdContext, build_spmd_federation  # noqa: E402
from fl4health_b200.utils.dataset import TensorDataset  # noqa: E402


def synthetic(n: int, seed: int) -> TensorDataset:
    gen = torch.Generator().manual_seed(seed)
    targets = torch.randint(0, 10, (n,), generator=gen)
    data = torch.randn(n, 3, 32, 32, generator=gen) * 0.5 + (targets.float().view(-1, 1, 1, 1) - 4.5) * 0.1
    return TensorDataset(data, targets)


def engine_for(mode: str) -> EngineOptions:
    if mode == "fp32-eager":
        return EngineOptions(arena=False, fused_optimizer=False, cuda_graphs=False)
    if mode == "fp32":
        return EngineOptions(cuda_graphs=True, channels_last=True, table_grads=True)
    if mode == "bf16":
        return EngineOptions(cuda_graphs=True, channels_last=True, amp_dtype=torch.bfloat16, master_weights=True)
    raise ValueError(mode)


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("--rounds", type=int, default=30)
    parser.add_argument("--modes", nargs="+", default=["fp32-eager", "fp32", "bf16"])
    parser.add_argument("--collectives", nargs="+", default=["auto"])
    parser.add_argument("--lr", type=float, default=0.003)
    args = parser.parse_args()
    ctx = SpmdContext()

    class Hooks:
        def get_data_loaders(self, config):  # noqa: ANN001, ANN202
            bs = int(config["batch_size"])
            return (BatchedTensorLoader(synthetic(2048, 100 + ctx.rank), bs, shuffle=True, drop_last=True, placement="device", device=self.device,
                                        generator=torch.Generator().manual_seed(7 + ctx.rank)),
                    BatchedTensorLoader(synthetic(512, 900 + ctx.rank), bs, placement="device", device=self.device))

        def get_criterion(self, config):  # noqa: ANN001, ANN202
            return nn.CrossEntropyLoss()

        def get_optimizer(self, config):  # noqa: ANN001, ANN202
            return torch.optim.SGD(self.model.parameters(), lr=args.lr, momentum=0.9)

    for mode in args.modes:
        for collectives in args.collectives:
            if mode == "fp32-eager":
                os.environ["FL4H_TC_CONV"] = "0"  # stock convolutions too
            else:
                os.environ.pop("FL4H_TC_CONV", None)
            os.environ["FL4H_COLLECTIVES"] = collectives
            ctx.collective_backend = collectives
            torch.manual_seed(1234 + ctx.rank)
            client, server = fl_variants.build("fedavg", Hooks, ctx, engine_for(mode), args.rounds, local_steps=8, batch_size=32)
            build_spmd_federation(ctx, server, client, fused=False if collectives == "nccl" else None)
            history, _ = server.fit(num_rounds=args.rounds)
            losses = [round(loss, 4) for _, loss in history.losses_distributed]
            accuracy = [round(float(v), 4) for _, v in history.metrics_distributed.get("val - prediction - accuracy", [])]
            if ctx.rank == 0:
                print("PARITY " + json.dumps({"mode": mode, "collectives": "fused" if (ctx.fused is not None and collectives != "nccl") else
                                              ("nccl" if ctx.world_size > 1 else "local"), "world": ctx.world_size, "rounds": args.rounds, "lr": args.lr,
                                              "final_val_loss": losses[-1], "final_val_accuracy": accuracy[-1] if accuracy else None,
                                              "val_loss_every_5": losses[4::5], "val_accuracy_every_5": accuracy[4::5]}), flush=True)
    ctx.shutdown()


if __name__ == "__main__":
    main()
