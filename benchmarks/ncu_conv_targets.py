"""Launch each conv-family kernel a few times on the ResNet-18 layer-1 / layer-3 shapes (ncu target).

    ncu --set full --clock-control none --import-source on -k regex:"conv_tap_gemm|conv_wgrad|stem_|bn_apply_presum" \
        -c 12 -o gpurun_out/ncu_conv python benchmarks/ncu_conv_targets.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from fl4health_b200.ops import conv  # noqa: E402


def cl(t):  # noqa: ANN001, ANN201
    return t.contiguous(memory_format=torch.channels_last)


def main() -> None:
    dev = "cuda"
    flush = torch.empty(64 << 20, device=dev)
    for dtype in (torch.float32,):
        for (n, cin, cout, h, r, s) in ((32, 64, 64, 32, 3, 1), (32, 256, 256, 8, 3, 1)):
            x = cl(torch.randn(n, cin, h, h, device=dev).to(dtype))
            w = cl((torch.randn(cout, cin, r, r, device=dev) / (cin * r * r) ** 0.5).to(dtype))
            dy = cl(torch.randn(n, cout, h // s, h // s, device=dev).to(dtype))
            stats = torch.zeros(2, cout, device=dev)
            for _ in range(2):  # first pass warms the L2 / instruction caches, ncu profiles every launch anyway
                flush.fill_(0.0)
                conv.conv2d_forward(x, w, s, 1, stats)
                conv.conv2d_dgrad(dy, w, (h, h), s, 1)
                conv.conv2d_wgrad(x, dy, r, s, 1)
        x = cl(torch.randn(32, 3, 32, 32, device=dev))
        w = cl(torch.randn(64, 3, 3, 3, device=dev))
        dy = cl(torch.randn(32, 64, 32, 32, device=dev))
        stats = torch.zeros(2, 64, device=dev)
        for _ in range(2):
            conv.stem_forward(x, w, stats)
            conv.stem_wgrad(x, dy)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
