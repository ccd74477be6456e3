"""What the mixins expect of the client they are combined with (parity: ``fl4health/mixins/core_protocols.py:15-107``,
there a ladder of ``typing.Protocol`` classes used for static typing only).

Here the same ladder is *checkable data*: every contract lists the attributes and methods it adds to its parents, and
``Contract.missing(obj)`` / ``Contract.require(obj)`` / ``isinstance(obj, SomeContract)`` verify an object (or a class)
against the accumulated list -- so a mixin stacked on the wrong base fails with the names that are absent instead of an
``AttributeError`` in the middle of a round.  The class names are the reference's, and the inheritance between them
(Ditto / MR-MTL extend the adaptive-drift contract, which extends the flexible client's) is preserved.
"""

from __future__ import annotations

from typing import Any


def _requirement(contract: str, name: str) -> Any:
    def required(self: Any, *args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError(f"{contract} requires the client to provide {name}()")

    required.__name__ = required.__qualname__ = name
    required.__doc__ = f"Required by {contract}; provided by the client class the mixin is combined with."
    return required


class _ContractType(type):
    def __new__(mcs, name: str, bases: tuple[type, ...], namespace: dict[str, Any]) -> _ContractType:
        # every required method is also an attribute of the contract class (introspection, ``help()``): a placeholder
        # that says who has to provide it.  Contracts are never base classes of clients, so nothing inherits these.
        for method in namespace.get("methods", ()):
            namespace.setdefault(method, _requirement(name, method))
        return super().__new__(mcs, name, bases, namespace)

    def __instancecheck__(cls, candidate: Any) -> bool:
        return not cls.missing(candidate)

    def __subclasscheck__(cls, candidate: type) -> bool:
        return type.__subclasscheck__(cls, candidate) or (isinstance(candidate, type) and not cls.missing_methods(candidate))


class Contract(metaclass=_ContractType):
    """``attributes`` exist on set-up instances only; ``methods`` exist on the class already."""

    attributes: tuple[str, ...] = ()
    methods: tuple[str, ...] = ()

    @classmethod
    def _collected(cls, kind: str) -> tuple[str, ...]:
        names: dict[str, None] = {}
        for level in reversed(cls.__mro__):
            names.update(dict.fromkeys(level.__dict__.get(kind, ())))
        return tuple(names)

    @classmethod
    def all_methods(cls) -> tuple[str, ...]:
        return cls._collected("methods")

    @classmethod
    def all_attributes(cls) -> tuple[str, ...]:
        return cls._collected("attributes")

    @classmethod
    def missing_methods(cls, candidate: Any) -> list[str]:
        return [name for name in cls.all_methods() if not callable(getattr(candidate, name, None))]

    @classmethod
    def missing(cls, candidate: Any) -> list[str]:
        absent = cls.missing_methods(candidate)
        if not isinstance(candidate, type):
            absent += [name for name in cls.all_attributes() if not hasattr(candidate, name)]
        return absent

    @classmethod
    def require(cls, candidate: Any) -> None:
        absent = cls.missing(candidate)
        if absent:
            owner = candidate.__name__ if isinstance(candidate, type) else type(candidate).__name__
            raise TypeError(f"Protocol requirements not met. {owner} lacks {', '.join(absent)} ({cls.__name__}).")


class NumPyClientMinimalProtocol(Contract):
    methods = ("get_parameters", "fit", "evaluate", "set_parameters", "update_after_train")


class FlexibleClientProtocolPreSetup(NumPyClientMinimalProtocol):
    attributes = ("device", "initialized")
    methods = (
        "setup_client", "get_model", "get_data_loaders", "get_optimizer", "get_criterion",
        "compute_loss_and_additional_losses",
    )


class FlexibleClientProtocol(FlexibleClientProtocolPreSetup):
    attributes = ("model", "optimizers")
    methods = (
        "initialize_all_model_weights", "update_before_train", "validate",
        # the per-model step family a personalised mixin drives once per model it owns
        "_compute_preds_and_losses", "_apply_backwards_on_losses_and_take_step", "_train_step_with_model_and_optimizer",
        "_val_step_with_model", "predict_with_model", "_transform_gradients_with_model",
        "transform_target", "transform_gradients", "compute_training_loss", "compute_evaluation_loss",
    )


class AdaptiveDriftConstrainedProtocol(FlexibleClientProtocol):
    """(``adaptive_drift_constrained.py:23-32``)"""

    attributes = (
        "loss_for_adaptation", "drift_penalty_tensors", "drift_penalty_weight", "penalty_loss_function", "parameter_exchanger",
    )
    methods = ("compute_penalty_loss", "setup_client_and_return_all_model_parameters")


class DittoPersonalizedProtocol(AdaptiveDriftConstrainedProtocol):
    """(``personalized/ditto.py:30-44``)"""

    attributes = ("global_model", "optimizer_keys")
    methods = ("get_global_model", "_copy_optimizer_with_new_params", "set_initial_global_tensors", "safe_global_model")


class MrMtlPersonalizedProtocol(AdaptiveDriftConstrainedProtocol):
    """(``personalized/mr_mtl.py:27-32``)"""

    attributes = ("initial_global_model", "initial_global_tensors")
    methods = ("get_global_model",)
