"""Second BASELINE.json configuration: BERT-base sequence classification (AG-News shapes: 4 classes, sequence 128) with
local AdamW and a FedAdam server, one client per GPU.  Same timing rules as bench.py (CUDA events, barrier +
synchronize on both sides, max over ranks, L2 flushed between rounds); prints one JSON line.

    python benchmarks/bert_fedopt_round.py --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 benchmarks/bert_fedopt_round.py --gpus 4
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")

import torch  # noqa: E402
from torch import nn  # noqa: E402

from fl4health_b200 import ops  # noqa: E402
from fl4health_b200.clients.basic_client import BasicClient  # noqa: E402
from fl4health_b200.common.typing import ndarrays_to_parameters  # noqa: E402
from fl4health_b200.engine.options import EngineOptions  # noqa: E402
from fl4health_b200.metrics import Accuracy  # noqa: E402
from fl4health_b200.metrics.metric_aggregation import evaluate_metrics_aggregation_fn, fit_metrics_aggregation_fn  # noqa: E402
from fl4health_b200.models.bert import BertConfig, BertForSequenceClassification  # noqa: E402
from fl4health_b200.parallel.arena import attach_arena  # noqa: E402
from fl4health_b200.parallel.spmd import SpmdContext, build_spmd_federation  # noqa: E402
from fl4health_b200.servers.base_server import FlServer  # noqa: E402
from fl4health_b200.servers.client_manager import SimpleClientManager  # noqa: E402
from fl4health_b200.strategies.fedopt import FedAdam  # noqa: E402


class DeviceDictLoader:
    """Device-resident token batches as ``({"input_ids", "attention_mask"}, labels)`` (what a tokenised AG-News shard
    looks like after collation)."""

    def __init__(self, n: int, seq_len: int, vocab: int, batch_size: int, shuffle: bool, device: torch.device, seed: int) -> None:
        gen = torch.Generator().manual_seed(seed)
        self.labels = torch.randint(0, 4, (n,), generator=gen).to(device)
        ids = torch.randint(10, vocab, (n, seq_len), generator=gen)
        ids[:, 1] = self.labels.cpu() + 1
        lengths = torch.randint(seq_len // 2, seq_len + 1, (n,), generator=gen)
        self.ids = ids.to(device)
        self.mask = (torch.arange(seq_len)[None, :] < lengths[:, None]).long().to(device)
        self.batch_size, self.shuffle, self.n = batch_size, shuffle, n
        self.dataset = list(range(n))  # only len() is used

    def __len__(self) -> int:
        return self.n // self.batch_size

    def __iter__(self):  # noqa: ANN204
        order = torch.randperm(self.n, device=self.ids.device) if self.shuffle else torch.arange(self.n, device=self.ids.device)
        for b in range(len(self)):
            idx = order[b * self.batch_size : (b + 1) * self.batch_size]
            yield {"input_ids": self.ids[idx], "attention_mask": self.mask[idx]}, self.labels[idx]


def main() -> None:
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--local-steps", type=int, default=4)
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--seq-len", type=int, default=128)
    p.add_argument("--val-batches", type=int, default=2)
    p.add_argument("--collectives", default="auto", choices=["auto", "nccl", "fused"])
    p.add_argument("--layers", type=int, default=12)
    p.add_argument("--no-graphs", action="store_true")
    p.add_argument("--stock-linear", action="store_true", help="nn.Linear + nn.GELU (cuBLAS) instead of the tcgen05 kernels")
    args = p.parse_args()

    os.environ["FL4H_COLLECTIVES"] = args.collectives
    ctx = SpmdContext()
    world, device = ctx.world_size, ctx.device
    assert world == args.gpus
    cfg = BertConfig(num_hidden_layers=args.layers, max_position_embeddings=max(args.seq_len, 128))
    if args.stock_linear:
        os.environ["FL4H_TC_DISABLE"] = "1"
    engine = EngineOptions(cuda_graphs=not args.no_graphs, amp_dtype=torch.bfloat16, master_weights=True)

    class BertClient(BasicClient):
        def get_model(self, config):  # noqa: ANN001, ANN202
            torch.manual_seed(99)
            return BertForSequenceClassification(cfg, 4)

        def get_data_loaders(self, config):  # noqa: ANN001, ANN202
            bs = int(config["batch_size"])
            return (DeviceDictLoader(64 * bs, args.seq_len, cfg.vocab_size, bs, True, self.device, 10 + ctx.rank),
                    DeviceDictLoader(args.val_batches * bs, args.seq_len, cfg.vocab_size, bs, False, self.device, 900 + ctx.rank))

        def get_criterion(self, config):  # noqa: ANN001, ANN202
            return nn.CrossEntropyLoss()

        def get_optimizer(self, config):  # noqa: ANN001, ANN202
            return torch.optim.AdamW(self.model.parameters(), lr=5e-5, weight_decay=0.01)

    def config_fn(server_round: int) -> dict:
        return {"current_server_round": server_round, "local_steps": args.local_steps, "batch_size": args.batch_size}

    client = BertClient(Path("."), [Accuracy()], device, client_name=f"rank{ctx.rank}", engine_options=engine)
    torch.manual_seed(99)
    template = BertForSequenceClassification(cfg, 4)
    n_params = sum(p.numel() for p in template.parameters())
    template_arena = attach_arena(template, device, with_grad=False)
    strategy = FedAdam(initial_parameters=ndarrays_to_parameters(template_arena.ndarrays()), eta=1e-3, tau=1e-6,
                       min_fit_clients=world, min_evaluate_clients=world, min_available_clients=world,
                       on_fit_config_fn=config_fn, on_evaluate_config_fn=config_fn,
                       fit_metrics_aggregation_fn=fit_metrics_aggregation_fn,
                       evaluate_metrics_aggregation_fn=evaluate_metrics_aggregation_fn)
    server = FlServer(SimpleClientManager(), {"n_server_rounds": args.steps + args.warmup}, strategy,
                      on_init_parameters_config_fn=config_fn, accept_failures=False)
    build_spmd_federation(ctx, server, client, fused=False if args.collectives == "nccl" else None)

    l2_flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=device)
    marks: dict = {}

    def on_round_end(server_round: int) -> None:
        l2_flush.fill_(1.0)
        if server_round == args.warmup:
            ctx.barrier()
            torch.cuda.synchronize()
            ops.reset_launch_count()
            marks["start"] = torch.cuda.Event(enable_timing=True)
            marks["start"].record()
        elif server_round == args.warmup + args.steps:
            marks["end"] = torch.cuda.Event(enable_timing=True)
            marks["end"].record()
            torch.cuda.synchronize()
            ctx.barrier()
            marks["launches"] = ops.launch_count()

    server.round_end_hooks = [on_round_end]
    history, _ = server.fit(num_rounds=args.warmup + args.steps)
    ms = ctx.all_reduce_max(marks["start"].elapsed_time(marks["end"])) / args.steps
    tokens = world * args.batch_size * args.seq_len * args.local_steps
    # forward+backward FLOPs of the encoder GEMMs + attention per token (6 x params-in-matmuls + attention scores)
    gemm_params = args.layers * (4 * cfg.hidden_size * cfg.hidden_size + 2 * cfg.hidden_size * cfg.intermediate_size)
    flops_per_token = 6 * gemm_params + 12 * args.layers * args.seq_len * cfg.hidden_size
    if ctx.rank == 0:
        print(json.dumps({
            "metric": "fl_client_rounds_per_sec_bert_base_fedadam", "value": world * 1000.0 / ms,
            "unit": "client-rounds/s (= FL rounds/s x N clients)", "federation_rounds_per_s": 1000.0 / ms, "ms_per_round": ms,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "dtype": "bf16", "data": "synthetic token ids, random-init BERT",
            "train_tokens_per_s": tokens / ms * 1e3, "train_tflops_per_gpu_in_round": flops_per_token * tokens / world / ms / 1e9,
            "config": {"model": f"bert-base-cased shapes ({n_params / 1e6:.1f}M params, {args.layers} layers)", "seq_len": args.seq_len,
                       "batch_per_client": args.batch_size, "local_steps": args.local_steps, "val_batches": args.val_batches,
                       "local_optimizer": "AdamW lr=5e-5", "strategy": "FedAdam eta=1e-3",
                       "linear": "nn.Linear (cuBLAS)" if args.stock_linear else "tcgen05 LinearAct (fused bias/GELU epilogue)",
                       "collectives": (("fused-nvls (multimem)" if ctx.fused.has_multicast else "fused-p2p") if ctx.fused is not None
                                       else ("nccl" if world > 1 else "local")),
                       "cuda_graphs": engine.cuda_graphs, "l2": "256 MiB write between rounds"},
            "gpu_launches": marks["launches"], "final_val_loss": history.losses_distributed[-1][1],
        }))
    ctx.shutdown()


if __name__ == "__main__":
    main()
