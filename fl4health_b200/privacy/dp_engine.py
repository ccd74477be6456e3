"""In-house DP-SGD engine (replaces the reference's dependency on Opacus; SURVEY §2.8 / hot-op L11).

Pieces, with the names users of the reference already know:

* ``GradSampleModule``  — wraps a model and computes *per-sample gradients* during the ordinary ``loss.backward()``
  with forward/backward hooks (Linear, Conv1d/2d, LayerNorm, GroupNorm, Embedding; anything else must be replaced by
  ``ModuleValidator.fix``, e.g. BatchNorm -> GroupNorm).  State-dict keys carry the ``_module.`` prefix, like Opacus.
  Two modes: ``"hooks"`` materialises ``p.grad_sample`` ([B, *p.shape], Opacus' contract); ``"ghost"`` (what
  ``PrivacyEngine.make_private`` uses) only *book-keeps*: per layer it adds the per-sample squared gradient norm to a
  [B] accumulator -- for Linear / Conv weights through the Gram identity ``||sum_t g_t a_t^T||^2 = sum_{t,s} (a_t.a_s)
  (g_t.g_s)`` whenever that is cheaper than forming the per-sample gradient -- and keeps the layer's (activation,
  back-propagated signal) pair; after the backward pass the optimizer turns the norms into clip factors and forms the
  CLIPPED SUM directly with one GEMM per layer (``(f . g)^T a``).  No [B, *p.shape] tensor for the large layers, every
  shape static: the whole clip / noise / step is CUDA-graph capturable.
* ``DPOptimizer``       — flat clipping: per-sample global L2 norm over all parameters, clip to ``max_grad_norm``, sum,
  add ``N(0, (noise_multiplier * max_grad_norm)^2)``, divide by the expected batch size, then the wrapped optimizer
  steps.  When the parameters' gradients are views of a flat arena the noise is ONE counter-RNG kernel over the flat
  gradient (``ops.flat.add_gaussian_``) and the wrapped optimizer can be the fused flat SGD/AdamW.
* ``DPDataLoader``      — Poisson sampling: each example joins a batch independently with probability
  ``batch_size / N`` (batches may be empty; the client skips those, as the reference does).
* ``PrivacyEngine.make_private`` — glue with the Opacus call signature used by ``InstanceLevelDpClient``.
"""

from __future__ import annotations

from collections.abc import Iterator
from logging import WARNING
from typing import Any

import torch
import torch.nn.functional as F
from torch import nn
from torch.optim import Optimizer

from fl4health_b200.common.logger import log
from fl4health_b200.ops import flat as flat_ops
from fl4health_b200.parallel.arena import arena_of

# ---------------------------------------------------------------------------------------------------------------
# per-sample gradient rules:  (module, activations, backprops) -> {parameter: grad_sample [B, *param.shape]}
# ---------------------------------------------------------------------------------------------------------------


def _linear_rule(layer: nn.Linear, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    out = {}
    if layer.weight.requires_grad:
        out[layer.weight] = torch.einsum("n...i,n...j->nij", g, a)
    if layer.bias is not None and layer.bias.requires_grad:
        out[layer.bias] = g.reshape(g.shape[0], -1, g.shape[-1]).sum(dim=1)
    return out


def _conv_rule(layer: nn.Module, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    one_d = isinstance(layer, nn.Conv1d)
    if one_d:  # treat as a 2-D convolution with height 1
        a, g = a.unsqueeze(-2), g.unsqueeze(-2)
        kernel, stride = (1, layer.kernel_size[0]), (1, layer.stride[0])
        padding = (0, layer.padding[0]) if not isinstance(layer.padding, str) else layer.padding
        dilation = (1, layer.dilation[0])
    else:
        kernel, stride, padding, dilation = layer.kernel_size, layer.stride, layer.padding, layer.dilation
    if isinstance(padding, str):
        raise NotImplementedError("string padding modes are not supported for per-sample gradients")
    n = a.shape[0]
    out = {}
    if layer.weight.requires_grad:
        cols = F.unfold(a, kernel, dilation=dilation, padding=padding, stride=stride)  # [B, Cin*k, L]
        gmat = g.reshape(n, layer.out_channels, -1)  # [B, Cout, L]
        groups = layer.groups
        cols = cols.reshape(n, groups, -1, cols.shape[-1])
        gmat = gmat.reshape(n, groups, layer.out_channels // groups, -1)
        grad = torch.einsum("ngol,ngil->ngoi", gmat, cols)
        out[layer.weight] = grad.reshape(n, *layer.weight.shape)
    if layer.bias is not None and layer.bias.requires_grad:
        out[layer.bias] = g.reshape(n, layer.out_channels, -1).sum(dim=2)
    return out


def _norm_rule(layer: nn.Module, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    out = {}
    if isinstance(layer, nn.LayerNorm):
        if layer.weight is not None and layer.weight.requires_grad:
            normed = F.layer_norm(a, layer.normalized_shape, eps=layer.eps)
            extra = tuple(range(1, a.dim() - len(layer.normalized_shape)))
            prod = normed * g
            out[layer.weight] = prod.sum(dim=extra) if extra else prod
        if layer.bias is not None and layer.bias.requires_grad:
            extra = tuple(range(1, a.dim() - len(layer.normalized_shape)))
            out[layer.bias] = g.sum(dim=extra) if extra else g
        return out
    assert isinstance(layer, nn.GroupNorm)
    if layer.weight is not None and layer.weight.requires_grad:
        normed = F.group_norm(a, layer.num_groups, eps=layer.eps)
        out[layer.weight] = torch.einsum("nc...,nc...->nc", normed, g)
    if layer.bias is not None and layer.bias.requires_grad:
        out[layer.bias] = g.reshape(g.shape[0], g.shape[1], -1).sum(dim=2)
    return out


def _embedding_rule(layer: nn.Embedding, a: torch.Tensor, g: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
    if not layer.weight.requires_grad:
        return {}
    n = a.shape[0]
    grad = torch.zeros(n, *layer.weight.shape, device=g.device, dtype=g.dtype)
    index = a.reshape(n, -1, 1).expand(-1, -1, layer.embedding_dim)
    grad.scatter_add_(1, index, g.reshape(n, -1, layer.embedding_dim))
    return {layer.weight: grad}


_RULES: list[tuple[tuple[type, ...], Any]] = [
    ((nn.Linear,), _linear_rule), ((nn.Conv1d, nn.Conv2d), _conv_rule), ((nn.LayerNorm, nn.GroupNorm), _norm_rule),
    ((nn.Embedding,), _embedding_rule),
]


def _rule_for(module: nn.Module) -> Any:
    for types, rule in _RULES:
        if isinstance(module, types):
            return rule
    return None


# ---------------------------------------------------------------------------------------------------------------
# book-keeping ("ghost") rules:  per-sample squared norms now, the clipped sum later
# ---------------------------------------------------------------------------------------------------------------
class _Deferred:
    """What a layer leaves behind in ghost mode: either per-sample gradients that were cheap to form (``samples``:
    parameter -> [B, *shape]) or the token-major factors of a weight gradient (``a``: [B, T, in], ``g``: [B, T, out])."""

    __slots__ = ("samples", "weight", "a", "g")

    def __init__(self) -> None:
        self.samples: dict[nn.Parameter, torch.Tensor] = {}
        self.weight: nn.Parameter | None = None
        self.a: torch.Tensor | None = None
        self.g: torch.Tensor | None = None

    def clipped_sums(self, factor: torch.Tensor) -> dict[nn.Parameter, torch.Tensor]:
        out = {p: torch.einsum("n,n...->...", factor.to(sample.dtype), sample) for p, sample in self.samples.items()}
        if self.weight is not None:
            assert self.a is not None and self.g is not None
            scaled = self.g * factor.to(self.g.dtype).view(-1, 1, 1)
            grad = scaled.reshape(-1, scaled.shape[-1]).t() @ self.a.reshape(-1, self.a.shape[-1])  # [out, in]: ONE GEMM
            out[self.weight] = grad.reshape(self.weight.shape)
        return out


def _factored_sq_norm(a: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """``||sum_t g[n,t] a[n,t]^T||_F^2`` per sample without forming the [out, in] products."""
    if a.shape[1] == 1:
        return a.pow(2).sum(dim=(1, 2)) * g.pow(2).sum(dim=(1, 2))
    return (torch.bmm(a, a.transpose(1, 2)) * torch.bmm(g, g.transpose(1, 2))).sum(dim=(1, 2))


def _gram_is_cheaper(tokens: int, fan_in: int, fan_out: int) -> bool:
    # Gram: B T^2 (in + out) multiply-adds and T^2 floats;  direct: B T in out and in*out floats per sample
    return tokens * (fan_in + fan_out) < fan_in * fan_out


def _ghost_rule(layer: nn.Module, a: torch.Tensor, g: torch.Tensor) -> tuple[torch.Tensor, _Deferred]:
    """(per-sample squared gradient norm of this layer's parameters [B], what to keep for the clipped sum)."""
    kept = _Deferred()
    factored: tuple[torch.Tensor, torch.Tensor] | None = None
    if isinstance(layer, nn.Linear) and layer.weight.requires_grad:
        a3, g3 = a.reshape(a.shape[0], -1, a.shape[-1]), g.reshape(g.shape[0], -1, g.shape[-1])
        if a3.shape[1] == 1 or _gram_is_cheaper(a3.shape[1], a3.shape[2], g3.shape[2]):
            factored = (a3, g3)
    elif isinstance(layer, nn.Conv2d) and layer.weight.requires_grad and layer.groups == 1 and not isinstance(layer.padding, str):
        tokens = g.shape[2] * g.shape[3]
        fan_in = layer.in_channels * layer.kernel_size[0] * layer.kernel_size[1]
        if _gram_is_cheaper(tokens, fan_in, layer.out_channels):
            cols = F.unfold(a, layer.kernel_size, dilation=layer.dilation, padding=layer.padding, stride=layer.stride)
            factored = (cols.transpose(1, 2), g.reshape(g.shape[0], layer.out_channels, -1).transpose(1, 2))
    if factored is None:
        kept.samples = _rule_for(layer)(layer, a, g)
    else:
        kept.weight, (kept.a, kept.g) = layer.weight, factored  # type: ignore[union-attr]
        bias = getattr(layer, "bias", None)
        if bias is not None and bias.requires_grad:
            kept.samples[bias] = factored[1].sum(dim=1)
    squared = sum(sample.reshape(sample.shape[0], -1).pow(2).sum(dim=1) for sample in kept.samples.values())
    if factored is not None:
        squared = squared + _factored_sq_norm(*factored)
    return squared, kept  # type: ignore[return-value]


# ---------------------------------------------------------------------------------------------------------------
# module validation / fixing
# ---------------------------------------------------------------------------------------------------------------
class ModuleValidator:
    """Finds layers whose per-sample gradients are undefined or unsupported (BatchNorm mixes samples) and replaces
    them with DP-compatible equivalents (BatchNorm -> GroupNorm with min(32, C) groups, as Opacus does)."""

    @staticmethod
    def validate(model: nn.Module, strict: bool = False) -> list[str]:
        errors = []
        for name, module in model.named_modules():
            if isinstance(module, nn.modules.batchnorm._BatchNorm):
                errors.append(f"{name}: BatchNorm cannot support training with differential privacy (mixes samples)")
            elif any(p.requires_grad for p in module.parameters(recurse=False)) and _rule_for(module) is None:
                errors.append(f"{name}: no per-sample gradient rule for {type(module).__name__}")
        if strict and errors:
            raise ValueError("\n".join(errors))
        return errors

    @staticmethod
    def is_valid(model: nn.Module) -> bool:
        return not ModuleValidator.validate(model)

    @staticmethod
    def fix(model: nn.Module) -> nn.Module:
        for name, child in list(model.named_children()):
            if isinstance(child, nn.modules.batchnorm._BatchNorm):
                groups = min(32, child.num_features)
                while child.num_features % groups != 0:
                    groups -= 1
                setattr(model, name, nn.GroupNorm(groups, child.num_features, affine=child.affine))
            else:
                ModuleValidator.fix(child)
        return model


# ---------------------------------------------------------------------------------------------------------------
# GradSampleModule
# ---------------------------------------------------------------------------------------------------------------
class GradSampleModule(nn.Module):
    def __init__(self, module: nn.Module, batch_first: bool = True, loss_reduction: str = "mean",
                 grad_sample_mode: str = "hooks") -> None:
        super().__init__()
        errors = [e for e in ModuleValidator.validate(module) if "BatchNorm" in e or "no per-sample" in e]
        if errors:
            raise ValueError("Model is not DP-compatible:\n" + "\n".join(errors))
        assert grad_sample_mode in ("hooks", "ghost"), "grad_sample_mode must be 'hooks' (materialise) or 'ghost' (book-keep)"
        self._module = module
        self.grad_sample_mode = grad_sample_mode
        self.sq_norms: torch.Tensor | None = None  # ghost mode: per-sample squared gradient norm over all layers so far
        self.deferred: list[_Deferred] = []
        self.loss_reduction = loss_reduction
        self.hooks_enabled = True
        self._handles: list[Any] = []
        self._activations: dict[nn.Module, list[torch.Tensor]] = {}
        self._add_hooks()

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        return self._module(*args, **kwargs)

    def _add_hooks(self) -> None:
        for module in self._module.modules():
            rule = _rule_for(module)
            if rule is None or not any(p.requires_grad for p in module.parameters(recurse=False)):
                continue
            self._handles.append(module.register_forward_hook(self._capture_activation))
            self._handles.append(module.register_full_backward_hook(self._capture_backprop))

    def remove_hooks(self) -> None:
        for handle in self._handles:
            handle.remove()
        self._handles = []

    def _capture_activation(self, module: nn.Module, inputs: tuple, output: Any) -> None:
        if not self.hooks_enabled or not module.training or not torch.is_grad_enabled():
            return
        self._activations.setdefault(module, []).append(inputs[0].detach())

    def _capture_backprop(self, module: nn.Module, grad_input: tuple, grad_output: tuple) -> None:
        if not self.hooks_enabled or module not in self._activations or not self._activations[module]:
            return
        activation = self._activations[module].pop()
        backprop = grad_output[0].detach()
        if self.loss_reduction == "mean":
            backprop = backprop * backprop.shape[0]  # undo the 1/B of the mean loss: per-sample grads are unscaled
        if self.grad_sample_mode == "ghost":
            squared, kept = _ghost_rule(module, activation, backprop)
            self.sq_norms = squared if self.sq_norms is None else self.sq_norms + squared
            self.deferred.append(kept)
            return
        for param, grad_sample in _rule_for(module)(module, activation, backprop).items():
            existing = getattr(param, "grad_sample", None)
            param.grad_sample = grad_sample if existing is None else existing + grad_sample  # type: ignore[attr-defined]

    def zero_grad(self, set_to_none: bool = True) -> None:
        for param in self._module.parameters():
            param.grad_sample = None  # type: ignore[attr-defined]
        self._activations = {}
        self.sq_norms, self.deferred = None, []
        super().zero_grad(set_to_none)

    def to_standard_module(self) -> nn.Module:
        self.remove_hooks()
        return self._module


def wrap_model(model: nn.Module, grad_sample_mode: str = "hooks", *args: Any, **kwargs: Any) -> GradSampleModule:
    return GradSampleModule(model, *args, grad_sample_mode=grad_sample_mode, **kwargs)


# ---------------------------------------------------------------------------------------------------------------
# DPOptimizer
# ---------------------------------------------------------------------------------------------------------------
class DPOptimizer(Optimizer):
    def __init__(self, optimizer: Optimizer, *, noise_multiplier: float, max_grad_norm: float,
                 expected_batch_size: int | None, loss_reduction: str = "mean", generator: torch.Generator | None = None,
                 module: nn.Module | None = None) -> None:
        self.original_optimizer = optimizer
        self.noise_multiplier = noise_multiplier
        self.max_grad_norm = max_grad_norm
        self.expected_batch_size = expected_batch_size
        self.loss_reduction = loss_reduction
        self.generator = generator
        self.param_groups = optimizer.param_groups
        self.defaults = optimizer.defaults
        self.state = optimizer.state
        self._step_count = 0
        self._module = module
        self._noise_state: torch.Tensor | None = None

    @property
    def params(self) -> list[nn.Parameter]:
        return [p for group in self.param_groups for p in group["params"] if p.requires_grad]

    def _assign(self, param: nn.Parameter, summed: torch.Tensor, first: bool) -> None:
        if param.grad is not None and first:
            param.grad.copy_(summed.to(param.grad.dtype))  # keeps arena gradient views alive
        elif param.grad is not None:
            param.grad.add_(summed.to(param.grad.dtype))
        else:
            param.grad = summed

    def _clip_and_accumulate_from_book(self, book: GradSampleModule) -> bool:
        """Ghost mode: norms were accumulated layer by layer; the clipped sum of every layer is formed now."""
        if book.sq_norms is None:
            return False
        clip_factor = (self.max_grad_norm / (book.sq_norms.sqrt() + 1e-6)).clamp(max=1.0)
        mine, seen = {id(p) for p in self.params}, set()
        for kept in book.deferred:
            for param, summed in kept.clipped_sums(clip_factor).items():
                if id(param) in mine:
                    self._assign(param, summed, first=id(param) not in seen)  # a shared layer contributes once per use
                    seen.add(id(param))
        book.sq_norms, book.deferred = None, []
        return True

    def clip_and_accumulate(self) -> bool:
        if isinstance(self._module, GradSampleModule) and self._module.grad_sample_mode == "ghost":
            return self._clip_and_accumulate_from_book(self._module)
        params = [p for p in self.params if getattr(p, "grad_sample", None) is not None]
        if not params:
            return False
        per_param_norms = [p.grad_sample.reshape(p.grad_sample.shape[0], -1).norm(2, dim=1) for p in params]  # type: ignore[attr-defined]
        per_sample_norms = torch.stack(per_param_norms, dim=1).norm(2, dim=1)
        clip_factor = (self.max_grad_norm / (per_sample_norms + 1e-6)).clamp(max=1.0)
        for p in params:
            summed = torch.einsum("i,i...", clip_factor.to(p.grad_sample.dtype), p.grad_sample)  # type: ignore[attr-defined]
            if p.grad is not None:
                p.grad.copy_(summed.to(p.grad.dtype))  # keeps arena gradient views alive
            else:
                p.grad = summed
            p.grad_sample = None  # type: ignore[attr-defined]
        return True

    def add_noise_and_scale(self) -> None:
        std = self.noise_multiplier * self.max_grad_norm
        scale = 1.0
        if self.loss_reduction == "mean" and self.expected_batch_size:
            scale = 1.0 / float(self.expected_batch_size)
        arena = arena_of(self._module) if self._module is not None else None
        inner = getattr(self._module, "_module", None)
        if arena is None and inner is not None:
            arena = arena_of(inner)
        params = self.params
        if (arena is not None and arena.grad is not None
                and all(p.grad is not None and p.grad.untyped_storage().data_ptr() == arena.grad.untyped_storage().data_ptr() for p in params)):
            # one kernel over the whole flat gradient: (g + std z) / B, the noise stream position lives on the device and
            # is advanced there, so a captured graph draws fresh noise on every replay
            if self._noise_state is None or self._noise_state.device != arena.grad.device:
                seed = int(torch.randint(0, 2**62, (1,), generator=self.generator).item())
                self._noise_state = flat_ops.make_noise_state(arena.grad.device, seed)
            flat_ops.add_gaussian_state_(arena.grad, std, self._noise_state, scale)
            return
        for p in params:
            if p.grad is None:
                continue
            if std > 0:  # default generator of the gradient's device: CUDA-graph safe (torch registers its offset)
                p.grad.add_(torch.randn(p.grad.shape, device=p.grad.device, dtype=p.grad.dtype), alpha=std)
            p.grad.mul_(scale)

    def pre_step(self) -> bool:
        if not self.clip_and_accumulate():
            return False
        self.add_noise_and_scale()
        return True

    def step(self, closure: Any = None) -> Any:  # type: ignore[override]
        loss = closure() if closure is not None else None
        if self.pre_step():
            self._step_count += 1
            return self.original_optimizer.step()
        return loss

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.params:
            p.grad_sample = None  # type: ignore[attr-defined]
        self.original_optimizer.zero_grad(set_to_none) if not _is_flat(self.original_optimizer) else self.original_optimizer.zero_grad()

    def state_dict(self) -> dict[str, Any]:  # type: ignore[override]
        return self.original_optimizer.state_dict()

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:  # type: ignore[override]
        self.original_optimizer.load_state_dict(state_dict)

    def __repr__(self) -> str:
        return f"DPOptimizer({self.original_optimizer!r}, sigma={self.noise_multiplier}, C={self.max_grad_norm})"


def _is_flat(optimizer: Optimizer) -> bool:
    from fl4health_b200.engine.fused_optim import _FlatOptimizer

    return isinstance(optimizer, _FlatOptimizer)


# ---------------------------------------------------------------------------------------------------------------
# Poisson data loader
# ---------------------------------------------------------------------------------------------------------------
class DPDataLoader:
    """Poisson-sampled batches over any loader that exposes ``dataset`` and ``batch_size``."""

    def __init__(self, dataset: Any, sample_rate: float, generator: torch.Generator | None = None, collate_fn: Any = None) -> None:
        self.dataset = dataset
        self.sample_rate = sample_rate
        self.generator = generator
        self.collate_fn = collate_fn
        self.batch_size = max(1, int(round(sample_rate * len(dataset))))

    @classmethod
    def from_data_loader(cls, data_loader: Any, generator: torch.Generator | None = None) -> DPDataLoader:
        n = len(data_loader.dataset)
        return cls(data_loader.dataset, sample_rate=data_loader.batch_size / n, generator=generator,
                   collate_fn=getattr(data_loader, "collate_fn", None))

    def __len__(self) -> int:
        return max(1, int(round(1.0 / self.sample_rate)))

    def __iter__(self) -> Iterator[Any]:
        n = len(self.dataset)
        for _ in range(len(self)):
            mask = torch.rand(n, generator=self.generator) < self.sample_rate
            indices = mask.nonzero(as_tuple=False).reshape(-1)
            get_batch = getattr(self.dataset, "get_batch", None)
            if get_batch is not None and indices.numel() > 0:
                yield get_batch(indices.to(self.dataset.data.device))
                continue
            samples = [self.dataset[int(i)] for i in indices]
            if not samples:
                first = self.dataset[0]
                yield torch.empty((0, *first[0].shape), dtype=first[0].dtype), torch.empty((0,), dtype=torch.as_tensor(first[1]).dtype)
                continue
            if self.collate_fn is not None:
                yield self.collate_fn(samples)
            else:
                yield torch.stack([s[0] for s in samples]), torch.stack([torch.as_tensor(s[1]) for s in samples])


# ---------------------------------------------------------------------------------------------------------------
# PrivacyEngine
# ---------------------------------------------------------------------------------------------------------------
class PrivacyEngine:
    def __init__(self, secure_mode: bool = False) -> None:
        if secure_mode:
            log(WARNING, "secure_mode (cryptographic RNG) is not implemented; using the default generator")

    def make_private(
        self, *, module: nn.Module, optimizer: Optimizer, data_loader: Any, noise_multiplier: float,
        max_grad_norm: float, batch_first: bool = True, loss_reduction: str = "mean", poisson_sampling: bool = True,
        clipping: str = "flat", noise_generator: torch.Generator | None = None, grad_sample_mode: str = "ghost",
        **kwargs: Any,
    ) -> tuple[GradSampleModule, DPOptimizer, Any]:
        assert clipping == "flat", "only flat clipping is implemented"
        wrapped = module if isinstance(module, GradSampleModule) else GradSampleModule(
            module, batch_first, loss_reduction, grad_sample_mode=grad_sample_mode)
        expected_batch_size = int(data_loader.batch_size) if getattr(data_loader, "batch_size", None) else None
        dp_loader = DPDataLoader.from_data_loader(data_loader, noise_generator) if poisson_sampling else data_loader
        dp_optimizer = DPOptimizer(optimizer, noise_multiplier=noise_multiplier, max_grad_norm=max_grad_norm,
                                   expected_batch_size=expected_batch_size, loss_reduction=loss_reduction,
                                   generator=noise_generator, module=wrapped)
        return wrapped, dp_optimizer, dp_loader
