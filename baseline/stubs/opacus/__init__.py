"""Import-only placeholders for the Opacus names FL4Health imports at module load; the non-DP benchmark path never
touches them."""

from torch import nn


class GradSampleModule(nn.Module):
    def __init__(self, *args, **kwargs) -> None:  # noqa: ANN002, ANN003
        raise NotImplementedError("opacus is not installed in this image (reference-arm placeholder)")


class PrivacyEngine:
    def __init__(self, *args, **kwargs) -> None:  # noqa: ANN002, ANN003
        raise NotImplementedError("opacus is not installed in this image (reference-arm placeholder)")
