"""Micro-benchmarks of the in-house BERT block kernels against the library composition they replace, at the shapes of
BASELINE config #4 (bert-base, batch 32, sequence 128, bf16): attention forward / backward vs
``scaled_dot_product_attention`` (flash), and LayerNorm(x + dropout(y)) forward / backward vs dropout + add + layer_norm.

CUDA events on the launching stream, warm-up, L2 evicted before every timed launch (a 256 MiB write), median of 30.
Prints one ``BLOCK {json}`` line per op.
"""

from __future__ import annotations

import json
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from fl4health_b200.ops.attention import packed_self_attention, packed_self_attention_reference  # noqa: E402
from fl4health_b200.ops.layer_norm import add_dropout_layer_norm, add_dropout_layer_norm_reference  # noqa: E402

FLUSH = None


def timed(fn, iters: int = 30) -> float:  # noqa: ANN001
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for _ in range(5):
        fn()
    samples = []
    for _ in range(iters):
        FLUSH.fill_(1.0)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        fn()
        end.record()
        torch.cuda.synchronize()
        samples.append(start.elapsed_time(end) * 1e3)
    return statistics.median(samples)


def bench_pair(name: str, ours, stock, inputs, flops: float | None = None, bytes_moved: float | None = None) -> None:  # noqa: ANN001
    record: dict = {"op": name}
    for label, fn in (("ours", ours), ("library", stock)):
        leaves = [t.detach().clone().requires_grad_(t.is_floating_point()) for t in inputs]
        out = fn(*leaves)
        upstream = torch.randn_like(out)
        record[f"{label}_fwd_us"] = round(timed(lambda: fn(*leaves)), 2)

        def fwd_bwd() -> None:
            for leaf in leaves:
                leaf.grad = None
            fn(*leaves).backward(upstream)

        record[f"{label}_fwd_bwd_us"] = round(timed(fwd_bwd), 2)
    if flops:
        record["ours_fwd_tflops"] = round(flops / record["ours_fwd_us"] * 1e-6, 1)
        record["ours_fwd_bwd_tflops"] = round(3.5 * flops / record["ours_fwd_bwd_us"] * 1e-6, 1)  # bwd: 5 GEMMs vs 2
    if bytes_moved:
        record["ours_fwd_gbps"] = round(bytes_moved / record["ours_fwd_us"] * 1e-3, 1)
    print("BLOCK " + json.dumps(record), flush=True)


def main() -> None:
    torch.manual_seed(0)
    batch, seq, heads, hidden = 32, 128, 12, 768
    qkv = (torch.randn(batch, seq, 3 * hidden, device="cuda")).to(torch.bfloat16)
    lengths = torch.randint(seq // 2, seq + 1, (batch,), device="cuda")
    mask = (torch.arange(seq, device="cuda")[None, :] < lengths[:, None]).to(torch.uint8)
    attention_flops = 4.0 * batch * heads * seq * seq * 64
    bench_pair("attention_b32_t128_h12", lambda x: packed_self_attention(x, mask, heads),
               lambda x: packed_self_attention_reference(x, mask, heads), [qkv], flops=attention_flops)
    bench_pair("attention_nomask", lambda x: packed_self_attention(x, None, heads),
               lambda x: packed_self_attention_reference(x, None, heads), [qkv], flops=attention_flops)
    rows = batch * seq
    y = torch.randn(rows, hidden, device="cuda").to(torch.bfloat16)
    res = torch.randn(rows, hidden, device="cuda").to(torch.bfloat16)
    weight, bias = torch.ones(hidden, device="cuda"), torch.zeros(hidden, device="cuda")
    for p in (0.0, 0.1):
        bench_pair(f"add_dropout_layernorm_p{p}", lambda a, b, w, c: add_dropout_layer_norm(a, b, w, c, 1e-12, p, True),
                   lambda a, b, w, c: add_dropout_layer_norm_reference(a.float(), b.float(), w, c, 1e-12, p, True).to(torch.bfloat16),  # = autocast
                   [y, res, weight, bias], bytes_moved=4.0 * rows * hidden * 2)


if __name__ == "__main__":
    main()
