"""Import-only placeholder: the reference's ``TorchMetric`` adapter type-checks against ``torchmetrics.Metric``; the
benchmark path never instantiates one."""


class Metric:
    def __init__(self, *args, **kwargs) -> None:  # noqa: ANN002, ANN003
        raise NotImplementedError("torchmetrics is not installed in this image (reference-arm placeholder)")
