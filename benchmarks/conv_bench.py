"""Per-layer timing of the tcgen05 convolution kernels against cuDNN (ResNet-18 / CIFAR shapes, batch 32).

CUDA-graph-captured loops of 20 launches, CUDA events, L2 kept warm (that is how the layers run inside a training step:
producer -> consumer within a few microseconds).  Prints one JSON line per (shape, dtype)."""

import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from fl4health_b200.ops import conv  # noqa: E402

SHAPES = [("l1_3x3", 32, 64, 64, 32, 3, 1), ("l2_3x3_s2", 32, 64, 128, 32, 3, 2), ("l2_1x1_s2", 32, 64, 128, 32, 1, 2),
          ("l2_3x3", 32, 128, 128, 16, 3, 1), ("l3_3x3_s2", 32, 128, 256, 16, 3, 2), ("l3_3x3", 32, 256, 256, 8, 3, 1),
          ("l4_3x3_s2", 32, 256, 512, 8, 3, 2), ("l4_3x3", 32, 512, 512, 4, 3, 1)]


def timed(fn, reps: int = 20) -> float:  # noqa: ANN001
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * reps) * 1e3  # microseconds per call


def main() -> None:
    torch.backends.cudnn.benchmark = True
    for dtype in (torch.float32, torch.bfloat16):
        for name, n, cin, cout, h, r, stride in SHAPES:
            pad = (r - 1) // 2
            x = torch.randn(n, cin, h, h, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
            w = (torch.randn(cout, cin, r, r, device="cuda") / (cin * r * r) ** 0.5).to(dtype).contiguous(memory_format=torch.channels_last)
            y = F.conv2d(x, w, None, stride, pad)
            dy = torch.randn_like(y).contiguous(memory_format=torch.channels_last)
            stats = torch.zeros(2, cout, device="cuda")
            flops = 2 * n * (h // stride) ** 2 * cout * cin * r * r
            row = {"shape": name, "dtype": "tf32" if dtype == torch.float32 else "bf16", "gflop": round(flops / 1e9, 3)}
            row["tc_fwd_us"] = round(timed(lambda: conv.conv2d_forward(x, w, stride, pad, stats)), 2)
            row["cudnn_fwd_us"] = round(timed(lambda: F.conv2d(x, w, None, stride, pad)), 2)
            row["tc_dgrad_us"] = round(timed(lambda: conv.conv2d_dgrad(dy, w, (h, h), stride, pad)), 2)
            row["cudnn_dgrad_us"] = round(timed(lambda: torch.ops.aten.convolution_backward(
                dy, x, w, None, [stride] * 2, [pad] * 2, [1, 1], False, [0, 0], 1, [True, False, False])), 2)
            row["tc_wgrad_us"] = round(timed(lambda: conv.conv2d_wgrad(x, dy, r, stride, pad)), 2)
            row["cudnn_wgrad_us"] = round(timed(lambda: torch.ops.aten.convolution_backward(
                dy, x, w, None, [stride] * 2, [pad] * 2, [1, 1], False, [0, 0], 1, [False, True, False])), 2)
            print("CONV " + json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
