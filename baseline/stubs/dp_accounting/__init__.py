"""Import-only placeholders for the ``dp_accounting`` names FL4Health imports at module load (the package is not in
this image).  Non-DP paths of the reference never touch them; anything that would compute a privacy guarantee raises."""


class _Unavailable:
    def __init__(self, *args, **kwargs) -> None:  # noqa: ANN002, ANN003
        raise NotImplementedError("dp_accounting is not installed in this image (reference-arm placeholder)")


class DpEvent(_Unavailable): ...


class DpEventBuilder(_Unavailable): ...


class GaussianDpEvent(DpEvent): ...


class PoissonSampledDpEvent(DpEvent): ...


class SampledWithoutReplacementDpEvent(DpEvent): ...


class SelfComposedDpEvent(DpEvent): ...
