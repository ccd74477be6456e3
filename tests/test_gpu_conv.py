"""tcgen05 implicit-GEMM convolution (``ops/conv.py`` / ``csrc/conv_tc.cu``) against a plain PyTorch fp32 reference."""

from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (name, N, Cin, Cout, H, R, stride) — the ResNet-18 / CIFAR layer shapes + a non-multiple batch
SHAPES = [
    ("l1_3x3", 32, 64, 64, 32, 3, 1),
    ("l2_3x3_s2", 32, 64, 128, 32, 3, 2),
    ("l2_1x1_s2", 32, 64, 128, 32, 1, 2),
    ("l2_3x3", 32, 128, 128, 16, 3, 1),
    ("l3_3x3_s2", 32, 128, 256, 16, 3, 2),
    ("l3_3x3", 32, 256, 256, 8, 3, 1),
    ("l4_3x3_s2", 32, 256, 512, 8, 3, 2),
    ("l4_3x3", 32, 512, 512, 4, 3, 1),
    ("l4_1x1_s2", 32, 256, 512, 8, 1, 2),
    ("ragged_batch", 5, 64, 64, 8, 3, 1),
]


def _rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    return float((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6))


def _make(n, cin, cout, h, r, dtype):  # noqa: ANN001, ANN202
    g = torch.Generator(device="cuda").manual_seed(n * 7 + cin + cout + h + r)
    x = torch.randn(n, cin, h, h, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, r, r, device="cuda", generator=g) / (cin * r * r) ** 0.5).to(dtype)
    w = w.contiguous(memory_format=torch.channels_last)
    return x, w


@pytest.fixture(autouse=True)
def _exact_reference():  # noqa: ANN202
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
def test_conv_forward_and_statistics(shape, dtype) -> None:  # noqa: ANN001
    from fl4health_b200.ops import conv

    _, n, cin, cout, h, r, stride = shape
    x, w = _make(n, cin, cout, h, r, dtype)
    assert conv.supported(x, w, stride, (r - 1) // 2)
    stats = torch.zeros(2, cout, device="cuda")
    y = conv.conv2d_forward(x, w, stride, (r - 1) // 2, stats)
    ref = F.conv2d(x.float(), w.float(), None, stride, (r - 1) // 2)
    tol = 4e-3 if dtype == torch.float32 else 2e-2
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert _rel_err(y, ref) < tol
    # the epilogue statistics are those of the STORED tensor
    yf = y.float()
    assert torch.allclose(stats[0], yf.sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2 * yf.abs().max().item())
    assert torch.allclose(stats[1], (yf * yf).sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["tf32", "bf16"])
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
def test_conv_backward(shape, dtype) -> None:  # noqa: ANN001
    from fl4health_b200.ops import conv

    _, n, cin, cout, h, r, stride = shape
    pad = (r - 1) // 2
    x, w = _make(n, cin, cout, h, r, dtype)
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, stride, pad)
    dy = torch.randn_like(ref).to(dtype).contiguous(memory_format=torch.channels_last)
    ref.backward(dy.float())
    dx = conv.conv2d_dgrad(dy, w, (h, h), stride, pad)
    dw = conv.conv2d_wgrad(x, dy, r, stride, pad)
    tol = 4e-3 if dtype == torch.float32 else 2e-2
    assert _rel_err(dx, xr.grad) < tol
    assert _rel_err(dw, wr.grad) < tol


def test_conv_autograd_matches_reference_end_to_end() -> None:
    from fl4health_b200.ops import conv

    x, w = _make(32, 64, 64, 16, 3, torch.float32)
    x1, w1 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    x2, w2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y1 = conv.conv2d(x1, w1, 1, 1)
    y2 = F.conv2d(x2, w2, None, 1, 1)
    (y1 * y1).sum().backward()
    (y2 * y2).sum().backward()
    assert _rel_err(y1, y2) < 4e-3 and _rel_err(x1.grad, x2.grad) < 8e-3 and _rel_err(w1.grad, w2.grad) < 8e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_stem_conv_forward_statistics_and_weight_gradient(dtype) -> None:  # noqa: ANN001
    from fl4health_b200.ops import conv

    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(32, 3, 32, 32, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 3, 3, 3, device="cuda", generator=g) / 27 ** 0.5).to(dtype).contiguous(memory_format=torch.channels_last)
    assert conv.stem_supported(x, w, 1, 1)
    stats = torch.zeros(2, 64, device="cuda")
    y = conv.stem_forward(x, w, stats)
    wr = w.float().requires_grad_(True)
    ref = F.conv2d(x.float(), wr, None, 1, 1)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert _rel_err(y, ref) < tol
    assert torch.allclose(stats[0], y.float().sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(stats[1], (y.float() ** 2).sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)
    dy = torch.randn_like(ref).to(dtype).contiguous(memory_format=torch.channels_last)
    ref.backward(dy.float())
    for _ in range(2):  # twice: the scratch accumulator must have reset itself
        dw = conv.stem_wgrad(x, dy)
        assert _rel_err(dw, wr.grad) < (1e-4 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("with_residual", [False, True])
def test_conv_epilogue_statistics_feed_batchnorm(dtype, with_residual: bool, monkeypatch) -> None:  # noqa: ANN001
    """``conv_bn_act``: convolution whose epilogue reduces the batch statistics + apply-only BatchNorm kernel, against
    the SAME convolution followed by the stock BatchNorm composition (``FL4H_BN_KERNEL=0``).  Both see bit-identical
    convolution outputs, so this isolates the statistics hand-off: outputs, input / weight / affine gradients and the
    running statistics."""
    from fl4health_b200.models.fused_layers import BatchNormAct2d, TcConv2d, conv_bn_act

    torch.manual_seed(3)
    conv = TcConv2d(64, 128, 3, stride=2, padding=1, bias=False).cuda().to(memory_format=torch.channels_last)
    bn = BatchNormAct2d(128, relu=True).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.2, 0.2)
    x0 = (torch.randn(32, 64, 16, 16, device="cuda") + 0.3).contiguous(memory_format=torch.channels_last)
    res0 = torch.randn(32, 128, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last) if with_residual else None
    if dtype == torch.bfloat16:
        conv = conv.to(torch.bfloat16)
        x0, res0 = x0.to(dtype), (res0.to(dtype) if res0 is not None else None)
    state = {k: v.clone() for k, v in bn.state_dict().items()}

    def run(kernel: bool):  # noqa: ANN202
        monkeypatch.setenv("FL4H_BN_KERNEL", "1" if kernel else "0")
        bn.load_state_dict(state)
        for p in (*conv.parameters(), *bn.parameters()):
            p.grad = None
        x = x0.clone().requires_grad_()
        res = res0.clone().requires_grad_() if res0 is not None else None
        y = conv_bn_act(conv, bn, x, res)
        (y.float() * torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y)).sum().backward()
        grads = [x.grad, conv.weight.grad, bn.weight.grad, bn.bias.grad] + ([res.grad] if res is not None else [])
        return y.detach().clone(), [g.detach().clone() for g in grads], {k: v.clone() for k, v in bn.state_dict().items()}

    y_ref, g_ref, s_ref = run(False)
    y_ker, g_ker, s_ker = run(True)
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-4  # bf16: the statistics come from the fp32 accumulators, the stock path re-reads the rounded output
    assert _rel_err(y_ker, y_ref) < tol
    for got, ref in zip(g_ker, g_ref):
        # an output within rounding distance of 0 may sit on the other side of the ReLU: single elements of a gradient
        # can flip (max-error is meaningless), the bulk must agree
        assert float((got.float() - ref.float()).abs().mean() / ref.float().abs().mean().clamp_min(1e-9)) < 5 * tol
    assert int(s_ker["num_batches_tracked"]) == int(s_ref["num_batches_tracked"]) == int(state["num_batches_tracked"]) + 1
    for key in ("running_mean", "running_var"):
        assert torch.allclose(s_ker[key], s_ref[key], rtol=tol, atol=tol * 0.1), key


def test_resnet18_step_launches_no_library_convolution() -> None:
    """One training step of the flagship model under the profiler: every convolution is one of ours."""
    from torch.profiler import ProfilerActivity, profile

    from fl4health_b200.models import resnet18_cifar

    model = resnet18_cifar().cuda().to(memory_format=torch.channels_last)
    x = torch.randn(32, 3, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
    target = torch.randint(0, 10, (32,), device="cuda")
    for _ in range(2):
        F.cross_entropy(model(x), target).backward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        F.cross_entropy(model(x), target).backward()
        torch.cuda.synchronize()
    names = [evt.key for evt in prof.key_averages()]
    ours = [n for n in names if "conv_tap_gemm_kernel" in n or "conv_wgrad_kernel" in n or "stem_" in n]
    library = [n for n in names if any(tag in n.lower() for tag in ("cudnn", "cutlass", "implicit_gemm", "xmma", "wgrad_alg"))
               and "sgemm" not in n.lower()]
    assert ours and not library, library
