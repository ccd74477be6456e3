class DPOptimizer:
    def __init__(self, *args, **kwargs) -> None:  # noqa: ANN002, ANN003
        raise NotImplementedError("opacus is not installed in this image (reference-arm placeholder)")
