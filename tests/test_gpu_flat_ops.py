"""Numerics of the sm_100a flat-arena kernels against plain PyTorch fp32 references (run on a B200)."""

import pytest
import torch

from fl4health_b200.ops import flat as F

pytestmark = pytest.mark.gpu
N = 1 << 20


def _rand(n=N, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(n, generator=g, device="cuda") * scale


def _hp(**kw):
    hp = F.make_hyper_params("cuda")
    for k, v in kw.items():
        hp[getattr(F, k)] = v
    return hp


def test_library_loaded():
    from fl4health_b200 import ops

    assert ops.available()


@pytest.mark.parametrize("anchor,cv,shadow,nesterov", [(False, False, False, False), (True, False, True, False),
                                                        (False, True, False, True), (True, True, True, False)])
def test_sgd_step_matches_reference(anchor, cv, shadow, nesterov):
    w, g = _rand(seed=1), _rand(seed=2, scale=0.1)
    a = _rand(seed=3) if anchor else None
    c = _rand(seed=4, scale=0.01) if cv else None
    results = []
    for use_kernel in (True, False):
        wk, mk = w.clone(), torch.zeros_like(w)
        sk = torch.zeros(N, dtype=torch.bfloat16, device="cuda") if shadow else None
        hp = _hp(HP_LR=0.05, HP_MOM=0.9, HP_WD=1e-4, HP_MU=0.1, HP_NESTEROV=float(nesterov), HP_FIRST=1.0)
        for _ in range(3):
            if use_kernel:
                F.sgd_step(wk, g, mk, hp, a, c, sk)
            else:
                F.sgd_step_reference(wk, g, mk, hp, a, c, sk)
        results.append((wk, mk, sk))
    torch.cuda.synchronize()
    assert torch.allclose(results[0][0], results[1][0], atol=1e-6, rtol=1e-5)
    assert torch.allclose(results[0][1], results[1][1], atol=1e-6, rtol=1e-5)
    if shadow:
        assert torch.allclose(results[0][2].float(), results[1][2].float(), atol=1e-2, rtol=1e-2)


def test_sgd_matches_torch_optim():
    w = _rand(seed=5)
    p = torch.nn.Parameter(w.clone())
    opt = torch.optim.SGD([p], lr=0.1, momentum=0.9, weight_decay=1e-3)
    wk, mk = w.clone(), torch.zeros_like(w)
    hp = _hp(HP_LR=0.1, HP_MOM=0.9, HP_WD=1e-3, HP_FIRST=1.0)
    for step in range(4):
        g = _rand(seed=10 + step, scale=0.1)
        p.grad = g.clone()
        opt.step()
        F.sgd_step(wk, g, mk, hp)
    assert torch.allclose(wk, p.data, atol=1e-6, rtol=1e-5)


def test_bf16_grad_path():
    w, g = _rand(seed=1), _rand(seed=2, scale=0.1).to(torch.bfloat16)
    wk, mk = w.clone(), torch.zeros_like(w)
    wr, mr = w.clone(), torch.zeros_like(w)
    F.sgd_step(wk, g, mk, _hp(HP_LR=0.05, HP_MOM=0.9, HP_FIRST=1.0))
    F.sgd_step_reference(wr, g, mr, _hp(HP_LR=0.05, HP_MOM=0.9, HP_FIRST=1.0))
    assert torch.allclose(wk, wr, atol=1e-6)


@pytest.mark.parametrize("decoupled", [True, False])
def test_adamw_matches_torch_optim(decoupled):
    w = _rand(seed=6)
    p = torch.nn.Parameter(w.clone())
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    opt = cls([p], lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2)
    wk, m, v = w.clone(), torch.zeros_like(w), torch.zeros_like(w)
    hp = _hp(HP_LR=1e-2, HP_B1=0.9, HP_B2=0.99, HP_EPS=1e-8, HP_WD=1e-2)
    for step in range(5):
        g = _rand(seed=20 + step, scale=0.1)
        p.grad = g.clone()
        opt.step()
        F.adamw_step(wk, g, m, v, hp, decoupled=decoupled)
    assert torch.allclose(wk, p.data, atol=2e-6, rtol=1e-4)


@pytest.mark.parametrize("k", [1, 2, 8, 16])
def test_weighted_sum(k):
    srcs = [_rand(seed=30 + i) for i in range(k)]
    coefs = [(i + 1) / sum(range(1, k + 1)) for i in range(k)]
    out = torch.empty(N, device="cuda")
    F.weighted_sum(out, srcs, coefs)
    ref = torch.zeros(N, device="cuda", dtype=torch.float64)
    for s, c in zip(srcs, coefs):
        ref += s.double() * c
    assert torch.allclose(out.double(), ref, atol=1e-5)
    # determinism: same inputs, same bits
    out2 = torch.empty(N, device="cuda")
    F.weighted_sum(out2, srcs, coefs)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("mode", [F.EPI_FEDADAM, F.EPI_FEDADAGRAD, F.EPI_FEDYOGI, F.EPI_SERVER_LR, F.EPI_MOMENTUM])
def test_weighted_sum_epilogues(mode):
    srcs = [_rand(seed=40 + i) for i in range(4)]
    coefs = [0.25] * 4
    cur = _rand(seed=50)
    kw = dict(mode=mode, eta=0.1, beta1=0.9, beta2=0.99, tau=1e-3, server_lr=0.7, momentum=0.9)
    outs = []
    for fn in (F.weighted_sum, F.weighted_sum_reference):
        m, v = _rand(seed=51, scale=0.1), _rand(seed=52, scale=0.1).abs()
        out = torch.empty(N, device="cuda")
        fn(out, srcs, coefs, current=cur.clone(), m=m, v=v, **kw)
        outs.append((out, m, v))
    for a, b in zip(outs[0], outs[1]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-4)


def test_bcast_unpack_and_scaffold():
    g, cs, cl = _rand(seed=60), _rand(seed=61), _rand(seed=62)
    w, a, cv = torch.empty(N, device="cuda"), torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    sh = torch.empty(N, device="cuda", dtype=torch.bfloat16)
    F.bcast_unpack(g, w, a, sh, cs, cl, cv)
    assert torch.equal(w, g) and torch.equal(a, g) and torch.equal(sh, g.to(torch.bfloat16))
    assert torch.allclose(cv, cs - cl)
    x, y = _rand(seed=63), _rand(seed=64)
    ci, dc = cl.clone(), torch.empty(N, device="cuda")
    F.scaffold_variate_update(x, y, cs, ci, dc, local_steps=5, lr=0.1)
    expect = cl - cs + (x - y) / 0.5
    assert torch.allclose(ci, expect, atol=1e-5) and torch.allclose(dc, expect - cl, atol=1e-5)


def test_reductions_and_clip():
    a, b = _rand(seed=70), _rand(seed=71)
    assert abs(F.sq_diff_sum(a, b).item() - ((a - b).double() ** 2).sum().item()) < 1e-2 * N ** 0.5
    assert abs(F.dot(a, b).item() - torch.dot(a.double(), b.double()).item()) < 1.0
    ga, gb = _rand(seed=72), _rand(seed=73)
    ref = torch.dot((a - b).double(), (0.3 * ga + 0.7 * gb).double()).item()
    assert abs(F.apfl_alpha_grad(a, b, ga, gb, 0.3).item() - ref) < 1.0
    x = a.clone()
    sq = F.sq_diff_sum(x)
    bit = torch.zeros(1, device="cuda")
    F.clip_scale_(x, sq, 10.0, bit)
    assert abs(x.norm().item() - 10.0) < 1e-2 and bit.item() == 0.0


def test_gaussian_noise_statistics():
    y = torch.zeros(N, device="cuda")
    F.add_gaussian_(y, 2.0, seed=123)
    assert abs(y.mean().item()) < 0.02 and abs(y.std().item() - 2.0) < 0.02
    y2 = torch.zeros(N, device="cuda")
    F.add_gaussian_(y2, 2.0, seed=123)
    assert torch.equal(y, y2)


def test_fedpm_vote():
    masks = [(torch.rand(4096, device="cuda") > 0.5).to(torch.uint8) for _ in range(3)]
    alpha, beta = torch.ones(4096, device="cuda"), torch.ones(4096, device="cuda")
    theta = F.fedpm_vote(masks, alpha, beta, bayesian=True)
    s = sum(m.float() for m in masks)
    assert torch.allclose(alpha, 1 + s) and torch.allclose(beta, 1 + 3 - s)
    assert torch.allclose(theta, (alpha - 1) / (alpha + beta - 2))


@pytest.mark.parametrize("n", [32, 1000 + 13, 1 << 20])
@pytest.mark.parametrize("k", [2, 8, 11])
def test_fedpm_packed_vote_matches_byte_vote(n: int, k: int):
    """Ballot-packed masks + popcount vote vs the plain PyTorch fp32 computation, incl. lengths that are not a multiple
    of 32 / of a warp's 1024 scores, fp32 and uint8 mask inputs, and the uniform-mean variant."""
    gen = torch.Generator(device="cuda").manual_seed(n + k)
    masks = [(torch.rand(n, device="cuda", generator=gen) < 0.2 + 0.05 * c).to(torch.uint8) for c in range(k)]
    words = torch.stack([F.pack_mask_bits(m if c % 2 else m.float()) for c, m in enumerate(masks)])
    assert words.shape == (k, (n + 31) // 32) and words.dtype == torch.int32
    for c, m in enumerate(masks):  # bit i % 32 of word i // 32
        assert torch.equal(F.unpack_mask_bits(words[c], n), m)
    votes = torch.stack(masks).float().sum(0)
    alpha, beta = torch.full((n,), 1.5, device="cuda"), torch.full((n,), 2.0, device="cuda")
    theta = F.fedpm_vote_packed(words, n, alpha, beta, bayesian=True)
    assert torch.equal(alpha, 1.5 + votes) and torch.equal(beta, 2.0 + k - votes)
    assert torch.allclose(theta, (alpha - 1) / (alpha + beta - 2), rtol=1e-6, atol=0)
    assert torch.allclose(F.fedpm_vote_packed(words, n, None, None, bayesian=False), votes / k, rtol=1e-6, atol=0)


def test_fused_moon_contrastive_matches_reference():
    from fl4health_b200.ops.contrastive import moon_contrastive, moon_contrastive_reference

    torch.manual_seed(0)
    z = torch.randn(32, 512, device="cuda", requires_grad=True)
    pos = torch.randn(32, 512, device="cuda", requires_grad=True)
    neg = torch.randn(3, 32, 512, device="cuda", requires_grad=True)
    ref = moon_contrastive_reference(z, pos, neg, 0.5)
    ref.backward()
    z2, p2, n2 = (t.detach().clone().requires_grad_() for t in (z, pos, neg))
    out = moon_contrastive(z2, p2, n2, 0.5)
    assert out.grad_fn is not None and "FusedMoon" in type(out.grad_fn).__name__
    out.backward()
    assert abs(out.item() - ref.item()) < 1e-5
    for a, b in ((z, z2), (pos, p2), (neg, n2)):
        assert torch.allclose(a.grad, b.grad, atol=1e-6, rtol=1e-4)


def test_fused_masked_parameter_statistics_and_gradient():
    from fl4health_b200.ops.masked import masked_parameter

    scores = torch.full((1 << 18,), 0.8, device="cuda", requires_grad=True)
    frozen = torch.full((1 << 18,), 2.0, device="cuda")
    out = masked_parameter(scores, frozen)
    keep = (out != 0).float().mean().item()
    p = torch.sigmoid(torch.tensor(0.8)).item()
    assert abs(keep - p) < 0.01 and set(out.unique().tolist()) <= {0.0, 2.0}
    out2 = masked_parameter(scores, frozen)
    assert not torch.equal(out, out2)  # fresh mask every call (device seed counter ticks)
    out.sum().backward()
    expected = 2.0 * p * p * (1 - p)
    assert torch.allclose(scores.grad, torch.full_like(scores, expected), atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("adam", [False, True])
def test_multi_tensor_step_matches_reference(adam: bool) -> None:
    """mt_optim.cu vs the per-slice PyTorch reference: mixed bf16/fp32 gradients, odd sizes, misaligned pointers."""
    from fl4health_b200.ops import multi_tensor as MT

    torch.manual_seed(0)
    dev = torch.device("cuda")
    sizes = [5, 4096, 4097, 33, 70001, 8, 12289]
    offsets, total = [], 0
    for n in sizes:
        offsets.append(total)
        total = (total + n + 31) // 32 * 32
    w = torch.randn(total, device=dev)
    m1, m2 = torch.randn(total, device=dev) * 0.1, torch.rand(total, device=dev) * 0.01
    anchor, cv = torch.randn(total, device=dev), torch.randn(total, device=dev) * 0.1
    shadow = torch.zeros(total, device=dev, dtype=torch.bfloat16)
    grads = []
    for i, n in enumerate(sizes):
        dtype = torch.bfloat16 if i % 2 == 0 else torch.float32
        store = torch.randn(n + 3, device=dev).to(dtype)
        grads.append(store[1:1 + n] if i in (2, 3) else store[:n])  # entries 2,3: misaligned base pointers
    hp = F.make_hyper_params(dev)
    for slot, value in ((F.HP_LR, 0.05), (F.HP_MOM, 0.9), (F.HP_WD, 1e-2), (F.HP_MU, 0.3), (F.HP_B1, 0.9), (F.HP_B2, 0.99),
                        (F.HP_EPS, 1e-8), (F.HP_STEP, 2.0), (F.HP_FIRST, 0.0), (F.HP_NESTEROV, 1.0)):
        hp[slot] = value
    ref = [t.detach().cpu().clone() for t in (w, m1, m2, anchor, cv, shadow, hp)]
    entries = [MT.TableEntry(g, o, n) for g, o, n in zip(grads, offsets, sizes)]
    MT.mt_step(entries, adam, w, m1, m2, hp, anchor, None if adam else cv, shadow)
    torch.cuda.synchronize()
    cpu_entries = [MT.TableEntry(g.detach().cpu(), o, n) for g, o, n in zip(grads, offsets, sizes)]
    rw, rm1, rm2, ranchor, rcv, rshadow, rhp = ref
    MT.mt_step_reference(cpu_entries, adam, rw, rm1, rm2, rhp, ranchor, None if adam else rcv, rshadow)
    for o, n in zip(offsets, sizes):
        assert torch.allclose(w[o:o + n].cpu(), rw[o:o + n], rtol=2e-5, atol=2e-6)
        assert torch.allclose(m1[o:o + n].cpu(), rm1[o:o + n], rtol=2e-5, atol=2e-6)
        assert torch.allclose(shadow[o:o + n].float().cpu(), rshadow[o:o + n].float(), rtol=1e-2, atol=1e-3)
        if adam:
            assert torch.allclose(m2[o:o + n].cpu(), rm2[o:o + n], rtol=2e-5, atol=1e-7)
    assert float(hp[F.HP_STEP]) == (3.0 if adam else 2.0) and float(hp[F.HP_FIRST]) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("relu,with_res", [(True, True), (True, False), (False, False), (False, True)])
@pytest.mark.parametrize("shape", [(32, 64, 32, 32), (8, 512, 4, 4), (5, 24, 7, 3), (64, 64, 32, 32)])
@pytest.mark.parametrize("fused", ["1", "0"])
def test_fused_batchnorm_act_matches_reference(dtype, relu, with_res, shape, fused, monkeypatch) -> None:
    """bn_act.cu (training fwd/bwd + eval fwd) vs F.batch_norm + add + relu in fp32; ``fused`` selects the grid-wide
    cooperative kernel or the two-kernel chain; the last shape exceeds the register-cached tile sizes."""
    from fl4health_b200.ops.bn_act import batch_norm_act, batch_norm_act_reference, kernel_eligible

    monkeypatch.setenv("FL4H_BN_FUSED", "0" if fused == "0" else "1")

    torch.manual_seed(1)
    dev = torch.device("cuda")
    n, c, h, w = shape
    x32 = (torch.randn(shape, device=dev) * 1.7 + 0.8).contiguous(memory_format=torch.channels_last)
    r32 = torch.randn(shape, device=dev).contiguous(memory_format=torch.channels_last) if with_res else None
    g32 = torch.randn(shape, device=dev).contiguous(memory_format=torch.channels_last)
    x = x32.to(dtype).requires_grad_(True)
    res = r32.to(dtype).requires_grad_(True) if with_res else None
    weight = (torch.rand(c, device=dev) + 0.5).requires_grad_(True)
    bias = torch.randn(c, device=dev).requires_grad_(True)
    rm, rv, nbt = torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5, torch.tensor(3, device=dev)
    assert kernel_eligible(x, res, 0.1, True, rm)
    # reference in fp32 on the (rounded) inputs
    xr = x.detach().float().requires_grad_(True)
    rr = res.detach().float().requires_grad_(True) if with_res else None
    wr, br = weight.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    from fl4health_b200 import ops

    before = ops.launch_count()
    y = batch_norm_act(x, weight, bias, rm, rv, nbt, True, 0.1, 1e-5, residual=res, relu=relu)
    y.backward(g32.to(dtype))
    torch.cuda.synchronize()
    assert ops.launch_count() - before == (4 if fused == "0" else 2)  # launches per fwd+bwd
    # Reference with the SAME ReLU mask as the kernel: pre-activations within rounding distance of 0 may land on
    # either side of the ReLU in two implementations that sum the batch statistics in different orders, and a single
    # flipped element would dominate the gradient comparison.
    pre = batch_norm_act_reference(xr, wr, br, rm_ref, rv_ref, True, 0.1, 1e-5, rr, False)
    y_ref = pre * (y.detach() > 0).float() if relu else pre
    y_ref.backward(g32.to(dtype).float())
    tol = dict(rtol=2e-2, atol=3e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-4)
    assert torch.allclose(y.float(), torch.relu(pre.detach()) if relu else pre.detach(), **tol)
    assert torch.allclose(rm, rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5)
    assert int(nbt) == 4
    m = n * h * w
    gtol = dict(rtol=3e-2, atol=3e-2 * (m ** 0.5) / 8) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=2e-3)
    assert torch.allclose(x.grad.float(), xr.grad, **tol)
    assert torch.allclose(weight.grad, wr.grad, **gtol), (weight.grad - wr.grad).abs().max()
    assert torch.allclose(bias.grad, br.grad, **gtol)
    if with_res:
        assert torch.allclose(res.grad.float(), rr.grad, **tol)
    # eval mode
    y_eval = batch_norm_act(x.detach(), weight.detach(), bias.detach(), rm, rv, nbt, False, 0.1, 1e-5,
                            residual=res.detach() if with_res else None, relu=relu)
    y_eval_ref = batch_norm_act_reference(xr.detach(), wr.detach(), br.detach(), rm_ref, rv_ref, False, 0.1, 1e-5,
                                          rr.detach() if with_res else None, relu)
    assert torch.allclose(y_eval.float(), y_eval_ref, **tol)
    assert int(nbt) == 4


@pytest.mark.gpu
@pytest.mark.parametrize("own_convs", [False, True])
def test_resnet_fused_bn_matches_stock_bn(own_convs: bool, monkeypatch) -> None:
    """Whole-model check: ResNet-18 with the fused BN path vs the same weights through the stock-op fallback.

    ``own_convs=False``: cuDNN convolutions in true fp32, so the two BatchNorm implementations are compared in
    isolation at tight tolerance.  ``own_convs=True``: the tcgen05 convolutions (TF32 products) whose epilogue feeds
    the batch statistics to the apply-only BatchNorm kernel -- the product path; the rounding of every convolution's
    inputs to 10 mantissa bits amplifies the last-bit differences of the statistics, so only direction, loss and
    running statistics are held to a tolerance there."""
    monkeypatch.setenv("FL4H_TC_CONV", "1" if own_convs else "0")
    import os

    from fl4health_b200.models import resnet18_cifar

    torch.manual_seed(0)
    dev = torch.device("cuda")
    # TF32 convolutions round their inputs to 10 mantissa bits: 1e-7 differences between the two BN paths would be
    # amplified to 1e-3 per layer and drown the comparison
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = False, False
    model = resnet18_cifar().to(dev).to(memory_format=torch.channels_last)
    x = torch.randn(16, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
    target = torch.randint(0, 10, (16,), device=dev)

    def run(fused: bool):
        model.zero_grad()
        state = {k: v.clone() for k, v in model.state_dict().items()}
        os.environ["FL4H_BN_KERNEL"] = "1" if fused else "0"  # 0: stock-op BatchNorm (and no statistics in the conv epilogue)
        try:
            loss = torch.nn.functional.cross_entropy(model(x), target)
            loss.backward()
        finally:
            os.environ.pop("FL4H_BN_KERNEL", None)
        grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        stats = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "tracked" in k}
        model.load_state_dict(state)
        return loss.item(), grads, stats

    try:
        l0, g0, s0 = run(False)
        l1, g1, s1 = run(True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert abs(l0 - l1) < (1e-3 if own_convs else 1e-4)
    # A pre-activation within rounding distance of 0 can land on different sides of a ReLU in the two implementations
    # (different summation order of the batch statistics); one such flip perturbs every upstream gradient at the
    # 1e-3 level and the nearest small layers by a few percent.  Direction and typical size must agree regardless.
    cosines = {n: float(torch.nn.functional.cosine_similarity(g0[n].flatten(), g1[n].flatten(), dim=0)) for n in g0}
    worst = sorted(cosines.items(), key=lambda kv: kv[1])[:5]
    assert worst[0][1] > 0.99, worst
    errors = sorted(float((g0[n] - g1[n]).abs().max() / g0[n].abs().max().clamp_min(1e-6)) for n in g0)
    assert errors[len(errors) // 2] < (0.15 if own_convs else 1e-2), errors
    for name in s0:
        assert torch.allclose(s0[name].float(), s1[name].float(), rtol=1e-3 if own_convs else 1e-4, atol=1e-4 if own_convs else 1e-5), name


@pytest.mark.gpu
@pytest.mark.parametrize("graphed", [False, True])
def test_overlapped_wgrad_matches_stock_conv_backward(graphed: bool, monkeypatch) -> None:
    """Conv2dOverlapWgrad (weight gradient on a side stream, joined at the end of backward) == nn.Conv2d backward,
    eagerly and as parallel branches of a captured CUDA graph."""
    from fl4health_b200.models import resnet18_cifar

    torch.manual_seed(0)
    dev = torch.device("cuda")
    model = resnet18_cifar().to(dev).to(memory_format=torch.channels_last).to(torch.bfloat16)
    x = torch.randn(16, 3, 32, 32, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    target = torch.randint(0, 10, (16,), device=dev)
    state = {k: v.clone() for k, v in model.state_dict().items()}

    def step() -> None:
        for p in model.parameters():
            p.grad = None
        torch.nn.functional.cross_entropy(model(x).float(), target).backward()

    def grads(overlap: str) -> dict[str, torch.Tensor]:
        monkeypatch.setenv("FL4H_OVERLAP_WGRAD", overlap)
        model.load_state_dict(state)
        if graphed and overlap == "1":
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
                    model.load_state_dict(state)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            captured = {n: p.grad for n, p in model.named_parameters()}
            model.load_state_dict(state)
            graph.replay()
            torch.cuda.synchronize()
            return {n: g.float().clone() for n, g in captured.items()}
        step()
        torch.cuda.synchronize()
        return {n: p.grad.float().clone() for n, p in model.named_parameters()}

    stock, overlapped = grads("0"), grads("1")
    for name in stock:
        scale = stock[name].abs().max().clamp_min(1e-6)
        assert float((stock[name] - overlapped[name]).abs().max() / scale) < 3e-2, name


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 384, 512), (32, 16, 512), (300, 200, 136), (4096, 1024, 1024)])
@pytest.mark.parametrize("relu,with_bias", [(False, False), (True, True)])
@pytest.mark.parametrize("variant", ["0", "1", "2", "3", "4"])
def test_tcgen05_linear_matches_reference(shape, relu, with_bias, variant, monkeypatch) -> None:
    """tc_gemm.cu (TMA -> tcgen05.mma -> TMEM -> epilogue) vs an fp32 reference, ragged tiles included; variants:
    one tile per CTA / persistent 128x128 / persistent 128x256 with double-buffered TMEM accumulators."""
    from fl4health_b200.ops.tc_gemm import kernel_eligible, linear_bias_act, linear_bias_act_reference

    monkeypatch.setenv("FL4H_TC_VARIANT", variant)

    m, n, k = shape
    if n % 8:
        pytest.skip("N must be a multiple of 8 for the kernel path")
    torch.manual_seed(m + n + k)
    dev = torch.device("cuda")
    x = torch.randn(m, k, device=dev).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(n, k, device=dev) / k ** 0.5).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(n, device=dev) if with_bias else None
    assert kernel_eligible(x, w, b)
    from fl4health_b200 import ops

    before = ops.launch_count()
    y = linear_bias_act(x, w, b, relu)
    torch.cuda.synchronize()
    assert ops.launch_count() - before == 1
    ref = linear_bias_act_reference(x.detach(), w.detach(), b, relu).float()
    assert torch.allclose(y.float(), ref, rtol=2e-2, atol=2e-2), (y.float() - ref).abs().max()
    # backward (library GEMMs + the kernel's ReLU mask)
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    pre = torch.nn.functional.linear(xr, wr, b)
    (pre * (y.detach() > 0).float() if relu else pre).backward(g.float())
    assert torch.allclose(x.grad.float(), xr.grad, rtol=3e-2, atol=3e-2)
    assert torch.allclose(w.grad.float(), wr.grad, rtol=3e-2, atol=3e-2 * (m ** 0.5))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(256, 384, 512), (300, 200, 136), (2048, 3072, 768)])
@pytest.mark.parametrize("variant", ["0", "1", "2", "3", "4"])
def test_tcgen05_linear_gelu_epilogue(shape, variant, monkeypatch) -> None:
    """GELU (erf) epilogue of tc_gemm.cu: the forward matches an fp32 reference, the pre-activation stored by the same
    epilogue feeds an exact GELU backward, and a no-grad call does not write the pre-activation at all."""
    from fl4health_b200.ops.tc_gemm import linear_bias_act, linear_bias_act_reference

    monkeypatch.setenv("FL4H_TC_VARIANT", variant)
    m, n, k = shape
    torch.manual_seed(m * 3 + n)
    dev = torch.device("cuda")
    x = torch.randn(m, k, device=dev).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(n, k, device=dev) / k ** 0.5).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(n, device=dev)
    y = linear_bias_act(x, w, b, "gelu")
    ref = linear_bias_act_reference(x.detach(), w.detach(), b, "gelu").float()
    assert torch.allclose(y.float(), ref, rtol=2e-2, atol=2e-2), (y.float() - ref).abs().max()
    with torch.no_grad():
        y_inf = linear_bias_act(x, w, b, "gelu")
    assert torch.equal(y_inf, y.detach())
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    torch.nn.functional.gelu(torch.nn.functional.linear(xr, wr, b)).backward(g.float())
    assert torch.allclose(x.grad.float(), xr.grad, rtol=4e-2, atol=4e-2), (x.grad.float() - xr.grad).abs().max()
    assert torch.allclose(w.grad.float(), wr.grad, rtol=4e-2, atol=4e-2 * (m ** 0.5))


@pytest.mark.gpu
def test_bert_layer_fused_matches_stock_modules(monkeypatch) -> None:
    """A BERT encoder whose projections run through the tcgen05 kernel (fused bias / GELU) vs the same weights through
    stock nn.Linear + nn.GELU in fp32: logits and every parameter gradient agree to bf16 accuracy."""
    from fl4health_b200 import ops
    from fl4health_b200.models.bert import BertConfig, BertForSequenceClassification

    monkeypatch.setenv("FL4H_TC_LINEAR", "always")  # default "auto" keeps mid-sized plain GEMMs on the library
    torch.manual_seed(7)
    dev = torch.device("cuda")
    cfg = BertConfig(vocab_size=512, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=1024,
                     max_position_embeddings=64, hidden_dropout_prob=0.0)
    model = BertForSequenceClassification(cfg, 8).to(dev)
    ref = BertForSequenceClassification(cfg, 8).to(dev)
    ref.load_state_dict(model.state_dict())
    ids = torch.randint(0, 512, (16, 64), device=dev)
    mask = (torch.arange(64, device=dev)[None, :] < torch.randint(32, 65, (16, 1), device=dev)).long()
    labels = torch.randint(0, 8, (16,), device=dev)
    model = model.to(torch.bfloat16)
    before = ops.launch_count()
    logits = model(ids, mask)
    # per layer: qkv, attn_out, ffn_in, ffn_out on the tcgen05 Linear kernel; + pooler + classifier; + one tcgen05 attention
    # launch per layer (bf16, T <= 128, head dim 64: measured 12 launches in total).  (The model was cast wholesale to bf16,
    # LayerNorm parameters included, so the fused LayerNorm kernel -- fp32 affine parameters -- stands aside.)
    linear_launches = 2 * 4 + 2
    attention_launches = (ops.launch_count() - before) - linear_launches
    assert attention_launches == cfg.num_hidden_layers, (ops.launch_count() - before, linear_launches)
    monkeypatch.setenv("FL4H_TC_LINEAR", "auto")
    before = ops.launch_count()
    with torch.no_grad():
        auto_logits = model(ids, mask)
    # auto: GEMMs this small stay on the library; what remains are the attention launches counted above
    assert ops.launch_count() - before == attention_launches
    assert torch.allclose(auto_logits.float(), logits.float(), rtol=5e-2, atol=5e-2)
    monkeypatch.setenv("FL4H_TC_LINEAR", "always")
    torch.nn.functional.cross_entropy(logits.float(), labels).backward()
    ref_logits = ref(ids, mask)
    torch.nn.functional.cross_entropy(ref_logits, labels).backward()
    assert torch.allclose(logits.float(), ref_logits, rtol=5e-2, atol=5e-2), (logits.float() - ref_logits).abs().max()
    for (name, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        cos = torch.nn.functional.cosine_similarity(p.grad.float().flatten(), q.grad.flatten(), dim=0)
        assert cos > 0.98, (name, float(cos))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["0", "1", "2", "3", "4"])
def test_tcgen05_linear_bf16_bias(variant, monkeypatch) -> None:
    """Master-weight mode hands the kernel a bf16 bias: the persistent variants read it as is (no cast kernel), the
    one-tile variant goes through an fp32 copy; both match the fp32 reference and produce a bf16 bias gradient."""
    from fl4health_b200.ops.tc_gemm import linear_bias_act, linear_bias_act_reference

    monkeypatch.setenv("FL4H_TC_VARIANT", variant)
    torch.manual_seed(5)
    dev = torch.device("cuda")
    x = torch.randn(512, 256, device=dev).to(torch.bfloat16)
    w = (torch.randn(384, 256, device=dev) / 16).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(384, device=dev).to(torch.bfloat16).requires_grad_(True)
    y = linear_bias_act(x, w, b, "gelu")
    ref = linear_bias_act_reference(x, w.detach(), b.detach(), "gelu").float()
    assert torch.allclose(y.float(), ref, rtol=2e-2, atol=2e-2)
    y.sum().backward()
    assert b.grad is not None and b.grad.dtype == torch.bfloat16 and w.grad.dtype == torch.bfloat16
    br = b.detach().float().requires_grad_(True)
    torch.nn.functional.gelu(torch.nn.functional.linear(x.float(), w.detach().float(), br)).sum().backward()
    assert torch.allclose(b.grad.float(), br.grad, rtol=5e-2, atol=0.5)


@pytest.mark.gpu
def test_optimizer_kernels_keep_infinite_parameters_finite_free_of_nan():
    """FedPM scores are legitimately +-inf after a Bayesian aggregate of exactly 0 or 1 (sigmoid_inverse).  With zero
    weight decay / drift weight the kernels must skip those terms like torch.optim does (0 * inf = NaN otherwise)."""
    dev = torch.device("cuda")
    w = torch.tensor([float("inf"), -float("inf"), 1.0, -2.0] * 8, device=dev)
    g = torch.full_like(w, 0.25)
    anchor = torch.zeros_like(w)
    hp = _hp(HP_LR=0.5)
    F.sgd_step(w, g, None, hp, anchor)
    assert not torch.isnan(w).any() and torch.isinf(w[0]) and torch.isinf(w[1])
    assert torch.allclose(w[2:4], torch.tensor([0.875, -2.125], device=dev))
    ref = torch.tensor([float("inf"), -float("inf"), 1.0, -2.0] * 8)
    F.sgd_step_reference(ref, g.cpu(), None, hp.cpu(), anchor.cpu())
    assert not torch.isnan(ref).any() and torch.allclose(ref[2:4], w[2:4].cpu())
    # table (multi-tensor) kernel, the path master-weight mode uses
    from fl4health_b200.ops.multi_tensor import TableEntry, mt_step

    master = torch.tensor([float("inf"), -float("inf"), 1.0, -2.0] * 8, device=dev)
    momentum = torch.zeros_like(master)
    grad = torch.full_like(master, 0.25)
    mt_step([TableEntry(grad, 0, master.numel())], False, master, momentum, None, _hp(HP_LR=0.5, HP_MOM=0.9, HP_FIRST=1.0),
            anchor=torch.zeros_like(master))
    assert not torch.isnan(master).any() and torch.isinf(master[0]) and torch.allclose(master[2:4], w[2:4])
