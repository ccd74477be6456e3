"""Python entry points for the flat-arena kernels (``csrc/flat_ops.cu``) with PyTorch reference fallbacks.

All functions operate on 1-D fp32 tensors whose length is a multiple of 4 (arenas are padded).  On CUDA tensors
the sm_100a kernels run on the current stream (CUDA-graph capturable); on CPU tensors the pure-PyTorch reference
runs — the same reference is the numerics oracle for the GPU tests.
"""

from __future__ import annotations

import ctypes
from collections.abc import Sequence

import torch

from fl4health_b200.ops import _lib

HP_LR, HP_MOM, HP_DAMP, HP_WD, HP_MU, HP_NESTEROV, HP_B1, HP_B2, HP_EPS, HP_STEP, HP_FIRST, HP_GSCALE = range(12)
HP_NOISE, HP_MAXNORM = 12, 13
HP_COUNT = 16

EPI_NONE, EPI_FEDADAM, EPI_FEDADAGRAD, EPI_FEDYOGI, EPI_SERVER_LR, EPI_MOMENTUM = range(6)
MAX_SRC = 16


def _use_kernel(t: torch.Tensor) -> bool:
    return t.is_cuda and _lib.load() is not None


def _check_flat(*tensors: torch.Tensor | None) -> int:
    n = -1
    for t in tensors:
        if t is None:
            continue
        assert t.dim() == 1 and t.is_contiguous(), "flat kernels take contiguous 1-D tensors"
        if n < 0:
            n = t.numel()
        assert t.numel() >= n
    assert n % 4 == 0, "arena length must be padded to a multiple of 4"
    return n


def make_hyper_params(device: torch.device | str) -> torch.Tensor:
    hp = torch.zeros(HP_COUNT, dtype=torch.float32, device=device)
    hp[HP_GSCALE] = 1.0
    return hp


# ---------------------------------------------------------------------------------------------------------------
# local optimizer steps
# ---------------------------------------------------------------------------------------------------------------
def sgd_step(
    w: torch.Tensor,
    grad: torch.Tensor,
    momentum_buf: torch.Tensor | None,
    hp: torch.Tensor,
    anchor: torch.Tensor | None = None,
    cv: torch.Tensor | None = None,
    shadow: torch.Tensor | None = None,
    more_ranges: bool = False,
) -> None:
    """w <- SGD(w, g + cv + mu (w - anchor) + wd w); also refreshes the bf16 compute shadow if given.  ``more_ranges``:
    further ranges of the same parameter group follow with the same ``hp`` block, whose first-step marker is then
    left armed for them."""
    n = _check_flat(w)
    if _use_kernel(w):
        lib = _lib.load(True)
        err = lib.fl4h_sgd_step(
            _lib.ptr(w), _lib.ptr(grad), _lib.ptr(momentum_buf), _lib.ptr(anchor), _lib.ptr(cv), _lib.ptr(shadow),
            _lib.ptr(hp), ctypes.c_int64(n), ctypes.c_int((1 if grad.dtype == torch.bfloat16 else 0) | (2 if more_ranges else 0)),
            _lib.stream_ptr(w.device),
        )
        _lib.check(err, "fl4h_sgd_step")
        _lib.count_launches(1 if more_ranges else 2)
        return
    sgd_step_reference(w, grad, momentum_buf, hp, anchor, cv, shadow, more_ranges)


def sgd_step_reference(w, grad, momentum_buf, hp, anchor=None, cv=None, shadow=None, more_ranges=False) -> None:  # noqa: ANN001
    h = hp.tolist()
    lr, mom, damp, wd, mu = h[HP_LR], h[HP_MOM], h[HP_DAMP], h[HP_WD], h[HP_MU]
    nesterov, first = h[HP_NESTEROV] != 0.0, h[HP_FIRST] != 0.0
    n = w.numel()
    g = grad[:n].to(torch.float32) * h[HP_GSCALE]
    if cv is not None:
        g = g + cv[:n]
    # zero coefficients are skipped (as torch.optim does): 0 * inf would be NaN and FedPM scores may be +-inf
    if anchor is not None and mu != 0.0:
        g = g + mu * (w - anchor[:n])
    if wd != 0.0:
        g = g + wd * w
    if mom != 0.0:
        assert momentum_buf is not None
        if first:
            momentum_buf[:n].copy_(g)
        else:
            momentum_buf[:n].mul_(mom).add_(g, alpha=1.0 - damp)
        upd = g + mom * momentum_buf[:n] if nesterov else momentum_buf[:n]
    else:
        upd = g
    w.sub_(lr * upd)
    if not more_ranges:
        hp[HP_FIRST] = 0.0
    if shadow is not None:
        shadow[:n].copy_(w)


def adamw_step(
    w: torch.Tensor,
    grad: torch.Tensor,
    exp_avg: torch.Tensor,
    exp_avg_sq: torch.Tensor,
    hp: torch.Tensor,
    anchor: torch.Tensor | None = None,
    shadow: torch.Tensor | None = None,
    decoupled: bool = True,
) -> None:
    n = _check_flat(w)
    if _use_kernel(w):
        lib = _lib.load(True)
        err = lib.fl4h_adamw_step(
            _lib.ptr(w), _lib.ptr(grad), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq), _lib.ptr(anchor),
            _lib.ptr(shadow), _lib.ptr(hp), ctypes.c_int64(n),
            ctypes.c_int(1 if grad.dtype == torch.bfloat16 else 0), ctypes.c_int(1 if decoupled else 0),
            _lib.stream_ptr(w.device),
        )
        _lib.check(err, "fl4h_adamw_step")
        _lib.count_launches(2)
        return
    adamw_step_reference(w, grad, exp_avg, exp_avg_sq, hp, anchor, shadow, decoupled)


def adamw_step_reference(w, grad, exp_avg, exp_avg_sq, hp, anchor=None, shadow=None, decoupled=True) -> None:  # noqa: ANN001
    hp[HP_STEP] += 1.0
    h = hp.tolist()
    lr, wd, mu, b1, b2, eps, step = h[HP_LR], h[HP_WD], h[HP_MU], h[HP_B1], h[HP_B2], h[HP_EPS], h[HP_STEP]
    n = w.numel()
    g = grad[:n].to(torch.float32) * h[HP_GSCALE]
    if anchor is not None and mu != 0.0:
        g = g + mu * (w - anchor[:n])
    if wd != 0.0:
        if decoupled:
            w.mul_(1.0 - lr * wd)
        else:
            g = g + wd * w
    exp_avg[:n].mul_(b1).add_(g, alpha=1.0 - b1)
    exp_avg_sq[:n].mul_(b2).addcmul_(g, g, value=1.0 - b2)
    bc1, bc2 = 1.0 - b1**step, 1.0 - b2**step
    denom = exp_avg_sq[:n].sqrt() / (bc2**0.5) + eps
    w.addcdiv_(exp_avg[:n], denom, value=-lr / bc1)
    if shadow is not None:
        shadow[:n].copy_(w)


# ---------------------------------------------------------------------------------------------------------------
# aggregation (+ server-optimizer epilogues)
# ---------------------------------------------------------------------------------------------------------------
def weighted_sum(
    out: torch.Tensor,
    srcs: Sequence[torch.Tensor],
    coefs: Sequence[float],
    *,
    mode: int = EPI_NONE,
    current: torch.Tensor | None = None,
    m: torch.Tensor | None = None,
    v: torch.Tensor | None = None,
    eta: float = 0.0,
    beta1: float = 0.0,
    beta2: float = 0.0,
    tau: float = 0.0,
    server_lr: float = 1.0,
    momentum: float = 0.0,
) -> torch.Tensor:
    """out = epilogue(sum_k coefs[k] * srcs[k]) in one pass, accumulating in the given (fixed) order."""
    n = _check_flat(out)
    assert 1 <= len(srcs) == len(coefs)
    if _use_kernel(out) and len(srcs) <= MAX_SRC:
        lib = _lib.load(True)
        k = len(srcs)
        ptrs = (ctypes.c_void_p * k)(*[s.data_ptr() for s in srcs])
        cf = (ctypes.c_float * k)(*[float(c) for c in coefs])
        err = lib.fl4h_weighted_sum(
            _lib.ptr(out), ptrs, cf, ctypes.c_int(k), _lib.ptr(current), _lib.ptr(m), _lib.ptr(v),
            ctypes.c_int(mode), ctypes.c_float(eta), ctypes.c_float(beta1), ctypes.c_float(beta2),
            ctypes.c_float(tau), ctypes.c_float(server_lr), ctypes.c_float(momentum), ctypes.c_int64(n),
            _lib.stream_ptr(out.device),
        )
        _lib.check(err, "fl4h_weighted_sum")
        _lib.count_launches(1)
        return out
    return weighted_sum_reference(
        out, srcs, coefs, mode=mode, current=current, m=m, v=v, eta=eta, beta1=beta1, beta2=beta2, tau=tau,
        server_lr=server_lr, momentum=momentum,
    )


def weighted_sum_reference(out, srcs, coefs, *, mode=EPI_NONE, current=None, m=None, v=None, eta=0.0, beta1=0.0,  # noqa: ANN001
                           beta2=0.0, tau=0.0, server_lr=1.0, momentum=0.0) -> torch.Tensor:
    n = out.numel()
    acc = torch.zeros_like(out)
    for src, c in zip(srcs, coefs):
        acc.add_(src[:n].to(out.dtype), alpha=float(c))
    if mode == EPI_NONE:
        out.copy_(acc)
        return out
    assert current is not None
    cur = current[:n]
    if mode in (EPI_FEDADAM, EPI_FEDADAGRAD, EPI_FEDYOGI):
        assert m is not None and v is not None
        d = acc - cur
        m[:n].mul_(beta1).add_(d, alpha=1.0 - beta1)
        d2 = d * d
        if mode == EPI_FEDADAM:
            v[:n].mul_(beta2).add_(d2, alpha=1.0 - beta2)
        elif mode == EPI_FEDADAGRAD:
            v[:n].add_(d2)
        else:
            v[:n].sub_((1.0 - beta2) * d2 * torch.sign(v[:n] - d2))
        out.copy_(cur + eta * m[:n] / (v[:n].sqrt() + tau))
    elif mode == EPI_SERVER_LR:
        out.copy_(cur + server_lr * (acc - cur))
    elif mode == EPI_MOMENTUM:
        assert m is not None
        m[:n].mul_(momentum).add_(acc)
        out.copy_(cur + server_lr * m[:n])
    else:
        raise ValueError(f"unknown epilogue mode {mode}")
    return out


# ---------------------------------------------------------------------------------------------------------------
# broadcast unpack / SCAFFOLD
# ---------------------------------------------------------------------------------------------------------------
def bcast_unpack(
    incoming: torch.Tensor,
    w: torch.Tensor | None = None,
    anchor: torch.Tensor | None = None,
    shadow: torch.Tensor | None = None,
    c_server: torch.Tensor | None = None,
    c_local: torch.Tensor | None = None,
    cv_out: torch.Tensor | None = None,
) -> None:
    """One read of the incoming global buffer -> w, FedProx anchor, bf16 shadow, SCAFFOLD (c - c_i)."""
    n = _check_flat(incoming)
    if _use_kernel(incoming):
        lib = _lib.load(True)
        err = lib.fl4h_bcast_unpack(
            _lib.ptr(incoming), _lib.ptr(w), _lib.ptr(anchor), _lib.ptr(shadow), _lib.ptr(c_server),
            _lib.ptr(c_local), _lib.ptr(cv_out), ctypes.c_int64(n), _lib.stream_ptr(incoming.device),
        )
        _lib.check(err, "fl4h_bcast_unpack")
        _lib.count_launches(1)
        return
    if w is not None:
        w[:n].copy_(incoming)
    if anchor is not None:
        anchor[:n].copy_(incoming)
    if shadow is not None:
        shadow[:n].copy_(incoming)
    if cv_out is not None:
        assert c_server is not None and c_local is not None
        torch.sub(c_server[:n], c_local[:n], out=cv_out[:n])


def scaffold_variate_update(
    x_global: torch.Tensor, y_local: torch.Tensor, c_server: torch.Tensor, c_local: torch.Tensor,
    delta_c: torch.Tensor, local_steps: int, lr: float,
) -> None:
    """c_i+ = c_i - c + (x - y)/(K lr); delta_c = c_i+ - c_i; c_i <- c_i+  (scaffold_client.py:229-261)."""
    n = _check_flat(x_global)
    inv = 1.0 / (local_steps * lr)
    if _use_kernel(x_global):
        lib = _lib.load(True)
        err = lib.fl4h_scaffold_variate(
            _lib.ptr(x_global), _lib.ptr(y_local), _lib.ptr(c_server), _lib.ptr(c_local), _lib.ptr(delta_c),
            ctypes.c_float(inv), ctypes.c_int64(n), _lib.stream_ptr(x_global.device),
        )
        _lib.check(err, "fl4h_scaffold_variate")
        _lib.count_launches(1)
        return
    new_ci = c_local[:n] - c_server[:n] + (x_global - y_local[:n]) * inv
    torch.sub(new_ci, c_local[:n], out=delta_c[:n])
    c_local[:n].copy_(new_ci)


# ---------------------------------------------------------------------------------------------------------------
# reductions / DP helpers
# ---------------------------------------------------------------------------------------------------------------
def sq_diff_sum(a: torch.Tensor, b: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """sum((a-b)^2) (or sum(a^2)) as a device scalar; accumulates into ``out`` when given."""
    n = _check_flat(a)
    if out is None:
        out = torch.zeros(1, dtype=torch.float32, device=a.device)
    if _use_kernel(a):
        lib = _lib.load(True)
        err = lib.fl4h_sq_diff_sum(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), ctypes.c_int64(n), _lib.stream_ptr(a.device))
        _lib.check(err, "fl4h_sq_diff_sum")
        _lib.count_launches(1)
        return out
    d = a if b is None else a - b[:n]
    out.add_((d.double() * d.double()).sum().float())
    return out


def dot(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    n = _check_flat(a)
    if out is None:
        out = torch.zeros(1, dtype=torch.float32, device=a.device)
    if _use_kernel(a):
        lib = _lib.load(True)
        err = lib.fl4h_dot(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), ctypes.c_int64(n), _lib.stream_ptr(a.device))
        _lib.check(err, "fl4h_dot")
        _lib.count_launches(1)
        return out
    out.add_(torch.dot(a.double(), b[:n].double()).float())
    return out


def apfl_alpha_grad(
    w_local: torch.Tensor, w_global: torch.Tensor, g_local: torch.Tensor, g_global: torch.Tensor, alpha: float
) -> torch.Tensor:
    """sum <w_loc - w_glob, alpha g_loc + (1-alpha) g_glob> in one pass (apfl_base.py:98-116 without host syncs)."""
    n = _check_flat(w_local)
    out = torch.zeros(1, dtype=torch.float32, device=w_local.device)
    if _use_kernel(w_local):
        lib = _lib.load(True)
        err = lib.fl4h_apfl_alpha_grad(
            _lib.ptr(w_local), _lib.ptr(w_global), _lib.ptr(g_local), _lib.ptr(g_global), ctypes.c_float(alpha),
            _lib.ptr(out), ctypes.c_int64(n), _lib.stream_ptr(w_local.device),
        )
        _lib.check(err, "fl4h_apfl_alpha_grad")
        _lib.count_launches(1)
        return out
    out.add_(torch.dot((w_local - w_global[:n]).double(), (alpha * g_local[:n] + (1 - alpha) * g_global[:n]).double()).float())
    return out


def clip_scale_(x: torch.Tensor, sq_norm: torch.Tensor, clip: float, bit: torch.Tensor | None = None) -> None:
    """x *= min(1, clip/sqrt(sq_norm)); bit <- (norm <= clip).  Norm stays on device (no host sync)."""
    n = _check_flat(x)
    if _use_kernel(x):
        lib = _lib.load(True)
        err = lib.fl4h_clip_scale(_lib.ptr(x), _lib.ptr(sq_norm), ctypes.c_float(clip), _lib.ptr(bit), ctypes.c_int64(n),
                                  _lib.stream_ptr(x.device))
        _lib.check(err, "fl4h_clip_scale")
        _lib.count_launches(1)
        return
    norm = sq_norm.sqrt()
    x.mul_(torch.clamp(clip / (norm + 1e-12), max=1.0))
    if bit is not None:
        bit.copy_((norm <= clip).float().reshape(bit.shape))


def add_gaussian_(y: torch.Tensor, stddev: float, seed: int) -> None:
    n = _check_flat(y)
    if stddev == 0.0:
        return
    if _use_kernel(y):
        lib = _lib.load(True)
        err = lib.fl4h_add_gaussian(_lib.ptr(y), ctypes.c_float(stddev), ctypes.c_uint64(seed & (2**64 - 1)),
                                    ctypes.c_int64(n), _lib.stream_ptr(y.device))
        _lib.check(err, "fl4h_add_gaussian")
        _lib.count_launches(1)
        return
    gen = torch.Generator(device=y.device).manual_seed(seed & (2**63 - 1))
    y.add_(torch.randn(y.shape, generator=gen, device=y.device, dtype=y.dtype), alpha=stddev)


def make_noise_state(device: torch.device | str, seed: int) -> torch.Tensor:
    """``[seed, draws]`` (int64) on the device: the position of a counter-based noise stream that kernels advance."""
    return torch.tensor([seed & (2**63 - 1), 0], dtype=torch.int64, device=device)


def add_gaussian_state_(y: torch.Tensor, stddev: float, state: torch.Tensor, scale: float = 1.0) -> None:
    """``y = (y + stddev * N(0, 1)) * scale`` drawing from the device-resident stream ``state`` (``make_noise_state``),
    which is advanced on the device: safe inside a captured CUDA graph (every replay draws new noise)."""
    n = _check_flat(y)
    if _use_kernel(y):
        lib = _lib.load(True)
        err = lib.fl4h_add_gaussian_state(_lib.ptr(y), ctypes.c_float(stddev), ctypes.c_float(scale), _lib.ptr(state),
                                          ctypes.c_int64(n), _lib.stream_ptr(y.device))
        _lib.check(err, "fl4h_add_gaussian_state")
        _lib.count_launches(2)
        return
    seed, draws = (int(v) for v in state.tolist())
    gen = torch.Generator(device=y.device).manual_seed((seed * 1_000_003 + draws) & (2**63 - 1))
    if stddev != 0.0:
        y.add_(torch.randn(y.shape, generator=gen, device=y.device, dtype=y.dtype), alpha=stddev)
    y.mul_(scale)
    state[1] += 1


def cast_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    n = _check_flat(src)
    if _use_kernel(src):
        lib = _lib.load(True)
        err = lib.fl4h_cast_bf16(_lib.ptr(src), _lib.ptr(dst), ctypes.c_int64(n), _lib.stream_ptr(src.device))
        _lib.check(err, "fl4h_cast_bf16")
        _lib.count_launches(1)
        return
    dst[:n].copy_(src)


def fedpm_vote(
    masks: Sequence[torch.Tensor], alpha: torch.Tensor | None, beta: torch.Tensor | None, bayesian: bool
) -> torch.Tensor:
    """FedPM mask vote (fl4health/strategies/fedpm.py:87-154): masks are uint8 {0,1} tensors of equal length."""
    n = masks[0].numel()
    theta = torch.empty(n, dtype=torch.float32, device=masks[0].device)
    k = len(masks)
    if _use_kernel(masks[0]) and k <= MAX_SRC:
        lib = _lib.load(True)
        ptrs = (ctypes.c_void_p * k)(*[mk.data_ptr() for mk in masks])
        err = lib.fl4h_fedpm_vote(ptrs, ctypes.c_int(k), _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(theta),
                                  ctypes.c_int(1 if bayesian else 0), ctypes.c_int64(n), _lib.stream_ptr(theta.device))
        _lib.check(err, "fl4h_fedpm_vote")
        _lib.count_launches(1)
        return theta
    s = torch.zeros(n, dtype=torch.float32, device=theta.device)
    for mk in masks:
        s.add_(mk.reshape(-1).float())
    if bayesian:
        assert alpha is not None and beta is not None
        alpha.add_(s)
        beta.add_(k - s)
        theta.copy_((alpha - 1.0) / (alpha + beta - 2.0))
    else:
        theta.copy_(s / k)
    return theta


def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """Flat {0,1} mask (uint8 / bool / float) -> ``ceil(n / 32)`` int32 words, score ``i`` at bit ``i % 32`` of word
    ``i // 32``.  What a FedPM client ships across GPUs: 1 bit per score."""
    flat = mask.reshape(-1)
    if flat.dtype == torch.bool:
        flat = flat.view(torch.uint8)
    elif flat.dtype not in (torch.uint8, torch.float32):
        flat = flat.to(torch.uint8)
    flat = flat.contiguous()
    n = flat.numel()
    n_words = (n + 31) // 32
    if _use_kernel(flat):
        words = torch.empty(n_words, dtype=torch.int32, device=flat.device)
        lib = _lib.load(True)
        err = lib.fl4h_pack_mask_bits(_lib.ptr(flat), ctypes.c_int(0 if flat.dtype == torch.uint8 else 1), _lib.ptr(words),
                                      ctypes.c_int64(n), _lib.stream_ptr(flat.device))
        _lib.check(err, "fl4h_pack_mask_bits")
        _lib.count_launches(1)
        return words
    bits = torch.zeros(n_words * 32, dtype=torch.int64, device=flat.device)
    bits[:n] = (flat != 0).to(torch.int64)
    packed = (bits.view(n_words, 32) << torch.arange(32, device=flat.device)).sum(dim=1)  # < 2**32, exact in int64
    return torch.where(packed >= 2**31, packed - 2**32, packed).to(torch.int32)


def unpack_mask_bits(words: torch.Tensor, n: int) -> torch.Tensor:
    """Inverse of ``pack_mask_bits`` (uint8, length ``n``); ``words`` may carry leading batch dimensions."""
    lanes = torch.arange(32, device=words.device)
    bits = (words.to(torch.int64).unsqueeze(-1) >> lanes) & 1
    return bits.reshape(*words.shape[:-1], -1)[..., :n].to(torch.uint8)


def fedpm_vote_packed(words: torch.Tensor, n: int, alpha: torch.Tensor | None, beta: torch.Tensor | None, bayesian: bool) -> torch.Tensor:
    """FedPM vote over ``K`` rows of bit-packed masks (``words``: ``[K, ceil(n/32)]`` int32, e.g. straight out of an
    all-gather).  Same update as ``fedpm_vote``: ``alpha += votes``, ``beta += K - votes``, posterior mode out."""
    assert words.dim() == 2 and words.dtype == torch.int32 and words.shape[1] * 32 >= n
    words = words.contiguous()
    k, n_words = words.shape
    theta = torch.empty(n, dtype=torch.float32, device=words.device)
    if _use_kernel(words):
        lib = _lib.load(True)
        err = lib.fl4h_fedpm_vote_packed(_lib.ptr(words), ctypes.c_int(k), ctypes.c_int64(n_words), _lib.ptr(alpha), _lib.ptr(beta),
                                         _lib.ptr(theta), ctypes.c_int(1 if bayesian else 0), ctypes.c_int64(n),
                                         _lib.stream_ptr(words.device))
        _lib.check(err, "fl4h_fedpm_vote_packed")
        _lib.count_launches(1)
        return theta
    votes = unpack_mask_bits(words, n).sum(dim=0).to(torch.float32)
    if bayesian:
        assert alpha is not None and beta is not None
        alpha.add_(votes)
        beta.add_(k - votes)
        theta.copy_((alpha - 1.0) / (alpha + beta - 2.0))
    else:
        theta.copy_(votes / k)
    return theta
