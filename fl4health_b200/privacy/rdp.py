"""Rényi-DP accounting primitives, implemented from the papers (no ``dp_accounting`` dependency).

* Gaussian mechanism:                      rdp(a) = a / (2 sigma^2)
* Poisson-subsampled Gaussian (add/remove one):   Mironov, Talwar, Zhang 2019 — exact binomial expansion for integer
  orders, the two-sided erfc series for fractional orders;
* fixed-size sampling without replacement (replace one): Wang, Balle, Kasiviswanathan 2019, Theorem 9 (ternary-|chi|^a
  bound with forward differences of the Gaussian CGF);
* RDP -> (eps, delta): the tightened conversion of Canonne, Kamath, Steinke 2020 (Prop. 12, arXiv:2004.00010 v4).

Validated against published tables in ``tests/test_privacy.py`` (reference oracles: SURVEY Appendix D).
"""

from __future__ import annotations

import math
from collections.abc import Sequence

import numpy as np
from scipy import special

# ---------------------------------------------------------------------------------------------------------------
# log-space helpers
# ---------------------------------------------------------------------------------------------------------------


def _log_add(logx: float, logy: float) -> float:
    a, b = min(logx, logy), max(logx, logy)
    if a == -np.inf:
        return b
    return math.log1p(math.exp(a - b)) + b


def _log_sub(logx: float, logy: float) -> float:
    if logx < logy:
        raise ValueError("The result of subtraction must be non-negative.")
    if logy == -np.inf:
        return logx
    if logx == logy:
        return -np.inf
    try:
        return math.log(math.expm1(logx - logy)) + logy
    except OverflowError:
        return logx


def _log_sub_sign(logx: float, logy: float) -> tuple[bool, float]:
    """log|e^logx - e^logy| and its sign (True = non-negative)."""
    if logx > logy:
        return True, _log_sub(logx, logy)
    if logx < logy:
        return False, _log_sub(logy, logx)
    return True, -np.inf


def _log_comb(n: float, k: float) -> float:
    return special.gammaln(n + 1) - special.gammaln(k + 1) - special.gammaln(n - k + 1)


def _log_erfc(x: float) -> float:
    return math.log(2) + special.log_ndtr(-x * 2**0.5)


# ---------------------------------------------------------------------------------------------------------------
# Poisson-subsampled Gaussian
# ---------------------------------------------------------------------------------------------------------------


def _log_a_int(q: float, sigma: float, alpha: int) -> float:
    log_a = -np.inf
    for i in range(alpha + 1):
        log_coef = _log_comb(alpha, i) + i * math.log(q) + (alpha - i) * math.log(1 - q)
        log_a = _log_add(log_a, log_coef + (i * i - i) / (2 * sigma**2))
    return float(log_a)


def _log_a_frac(q: float, sigma: float, alpha: float) -> float:
    log_a0, log_a1 = -np.inf, -np.inf
    z0 = sigma**2 * math.log(1 / q - 1) + 0.5
    i = 0
    while True:
        coef = special.binom(alpha, i)
        log_coef = math.log(abs(coef))
        j = alpha - i
        log_t0 = log_coef + i * math.log(q) + j * math.log(1 - q)
        log_t1 = log_coef + j * math.log(q) + i * math.log(1 - q)
        log_e0 = math.log(0.5) + _log_erfc((i - z0) / (math.sqrt(2) * sigma))
        log_e1 = math.log(0.5) + _log_erfc((z0 - j) / (math.sqrt(2) * sigma))
        log_s0 = log_t0 + (i * i - i) / (2 * sigma**2) + log_e0
        log_s1 = log_t1 + (j * j - j) / (2 * sigma**2) + log_e1
        if coef > 0:
            log_a0, log_a1 = _log_add(log_a0, log_s0), _log_add(log_a1, log_s1)
        else:
            log_a0, log_a1 = _log_sub(log_a0, log_s0), _log_sub(log_a1, log_s1)
        i += 1
        if max(log_s0, log_s1) < -30:
            break
    return _log_add(log_a0, log_a1)


def rdp_poisson_subsampled_gaussian(q: float, sigma: float, orders: Sequence[float]) -> np.ndarray:
    def one(alpha: float) -> float:
        if q == 0:
            return 0.0
        if sigma == 0:
            return np.inf
        if q == 1.0:
            return alpha / (2 * sigma**2)
        if np.isinf(alpha):
            return np.inf
        log_a = _log_a_int(q, sigma, int(alpha)) if float(alpha).is_integer() else _log_a_frac(q, sigma, alpha)
        return log_a / (alpha - 1)

    return np.array([one(a) for a in orders])


# ---------------------------------------------------------------------------------------------------------------
# sampling without replacement (Wang et al. 2019, Thm 9)
# ---------------------------------------------------------------------------------------------------------------


def _stable_inplace_diff_in_log(vec: np.ndarray, signs: np.ndarray, n: int) -> None:
    """One forward-difference pass over the first n+1 entries of a signed log-domain vector."""
    for j in range(n):
        if signs[j] == signs[j + 1]:  # same sign: |e^b - e^a|
            nonneg, value = _log_sub_sign(vec[j + 1], vec[j])
            signs[j] = nonneg if signs[j + 1] else not nonneg
            vec[j] = value
        else:  # opposite signs: magnitudes add, sign follows the later entry
            vec[j] = _log_add(vec[j], vec[j + 1])
            signs[j] = signs[j + 1]


def _forward_diffs(fun, n: int) -> np.ndarray:  # noqa: ANN001
    """log|Delta^k exp(fun)(-1)| for k = 1..n+2  (index k-1)."""
    vec = np.array([fun(float(i - 1)) for i in range(n + 3)])
    vec[0] = 0.0
    signs = np.ones(n + 3, dtype=bool)
    deltas = np.zeros(n + 2)
    for i in range(n + 2):
        _stable_inplace_diff_in_log(vec, signs, n + 2 - i)
        deltas[i] = vec[0]
    return deltas


def _rdp_sample_wor_gaussian_int(q: float, sigma: float, alpha: int) -> float:
    max_alpha = 256
    if alpha == 1:
        return 0.0

    def cgf(x: float) -> float:  # (x) * rdp_gauss(x + 1)
        return x * (x + 1) / (2.0 * sigma**2)

    def func(x: float) -> float:  # (x - 1) * rdp_gauss(x)
        return x * (x - 1) / (2.0 * sigma**2)

    log_f2m1 = func(2.0) + math.log(1 - math.exp(-func(2.0)))
    deltas = _forward_diffs(cgf, alpha) if alpha <= max_alpha else None
    log_a = 0.0
    for i in range(2, alpha + 1):
        if i == 2:
            s = 2 * math.log(q) + _log_comb(alpha, 2) + min(math.log(4) + log_f2m1, func(2.0) + math.log(2))
        else:
            s = math.log(2) + cgf(i - 1)
            if deltas is not None:
                delta_lo = deltas[int(2 * math.floor(i / 2.0)) - 1]
                delta_hi = deltas[int(2 * math.ceil(i / 2.0)) - 1]
                s = min(s, math.log(4) + 0.5 * (delta_lo + delta_hi))
            s += i * math.log(q) + _log_comb(alpha, i)
        log_a = _log_add(log_a, s)
    return float(log_a) / (alpha - 1)


def rdp_sample_wor_gaussian(q: float, sigma: float, orders: Sequence[float]) -> np.ndarray:
    def one(alpha: float) -> float:
        if q == 0:
            return 0.0
        if sigma == 0 or np.isinf(alpha):
            return np.inf
        if q == 1.0:
            return alpha / (2 * sigma**2)
        if float(alpha).is_integer():
            return _rdp_sample_wor_gaussian_int(q, sigma, int(alpha))
        lo, hi = int(math.floor(alpha)), int(math.ceil(alpha))  # convexity of (a-1) rdp(a): interpolate
        t = alpha - lo
        return ((1 - t) * (lo - 1) * _rdp_sample_wor_gaussian_int(q, sigma, lo)
                + t * (hi - 1) * _rdp_sample_wor_gaussian_int(q, sigma, hi)) / (alpha - 1)

    return np.array([one(a) for a in orders])


def rdp_gaussian(sigma: float, orders: Sequence[float]) -> np.ndarray:
    return np.array([np.inf if sigma == 0 else a / (2 * sigma**2) for a in orders])


# ---------------------------------------------------------------------------------------------------------------
# conversions
# ---------------------------------------------------------------------------------------------------------------


def epsilon_from_rdp(orders: Sequence[float], rdp: Sequence[float], delta: float) -> float:
    if delta < 0:
        raise ValueError(f"Delta cannot be negative. Found {delta}.")
    if delta == 0:
        return 0.0 if all(r == 0 for r in rdp) else np.inf
    best = np.inf
    for a, r in zip(orders, rdp):
        if a < 1 or r < 0:
            raise ValueError("orders must be >= 1 and rdp >= 0")
        if delta**2 + math.expm1(-r) > 0:
            eps = 0.0  # delta <= sqrt(1 - exp(-KL)) already holds
        elif a > 1.01:
            eps = r + math.log1p(-1 / a) - math.log(delta * a) / (a - 1)
        else:
            eps = np.inf
        best = min(best, eps)
    return max(0.0, float(best))


def delta_from_rdp(orders: Sequence[float], rdp: Sequence[float], epsilon: float) -> float:
    if epsilon < 0:
        raise ValueError(f"Epsilon cannot be negative. Found {epsilon}.")
    best = np.inf
    for a, r in zip(orders, rdp):
        if a < 1 or r < 0:
            raise ValueError("orders must be >= 1 and rdp >= 0")
        log_delta = -np.inf if r == 0 else 0.5 * math.log1p(-math.exp(-r))
        if a > 1.01:
            log_delta = min(log_delta, (a - 1) * (r - epsilon + math.log1p(-1 / a)) - math.log(a))
        best = min(best, log_delta)
    return min(math.exp(best), 1.0)
