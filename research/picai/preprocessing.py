"""Device-side preprocessing of multi-sequence prostate MRI cases (T2W / ADC / HBV + lesion annotation).

The reference pipeline (``research/picai/data/preprocessing.py:18-384``) works on SimpleITK images on the CPU, one
case at a time.  Here a scan is a ``Volume`` — a ``[D, H, W]`` tensor with its voxel spacing and physical origin —
and every transform is a tensor op (trilinear / nearest ``grid_sample`` for resampling, slicing / ``F.pad`` for the
crop-or-pad), so a whole study can be preprocessed on the GPU that trains on it; file IO stays at the edges
(``.pt`` dictionaries, or NIfTI/MHA through an optional reader supplied by the caller).

Transform order used by ``default_transforms`` (same as the reference's ``preprocess``): resample every sequence onto
the grid of the first scan → align origin → resample to the target spacing → centre crop / pad to the target size →
binarise the annotation.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Sequence
from dataclasses import dataclass, field, replace
from pathlib import Path

import torch
import torch.nn.functional as F

Triple = tuple[float, float, float]


@dataclass
class Volume:
    """A 3-D scan: ``data[D, H, W]``, voxel ``spacing`` in mm (depth, height, width), ``origin`` = physical position of
    voxel (0, 0, 0) in mm.  Axes are assumed aligned with the patient axes (direction = identity), which is what the
    PI-CAI archive provides after ``AlignOriginAndDirection``."""

    data: torch.Tensor
    spacing: Triple = (1.0, 1.0, 1.0)
    origin: Triple = (0.0, 0.0, 0.0)

    @property
    def size(self) -> tuple[int, int, int]:
        d, h, w = self.data.shape
        return int(d), int(h), int(w)

    @property
    def physical_size(self) -> Triple:
        return tuple(n * s for n, s in zip(self.size, self.spacing))  # type: ignore[return-value]


@dataclass
class PreprocessingSettings:
    """Target geometry.  Any two of ``size`` / ``spacing`` / ``physical_size`` determine the third; if all three are
    given they must agree (parity: ``preprocessing.py:18-60``)."""

    scans_write_dir: Path
    annotation_write_dir: Path
    size: tuple[int, int, int] | None = None
    physical_size: Triple | None = None
    spacing: Triple | None = None

    def __post_init__(self) -> None:
        if self.size is not None and self.spacing is not None:
            implied = tuple(n * s for n, s in zip(self.size, self.spacing))
            if self.physical_size is None:
                self.physical_size = implied  # type: ignore[assignment]
            else:
                assert all(abs(a - b) < 1e-6 for a, b in zip(implied, self.physical_size)), "size x spacing must equal physical_size"
        elif self.size is None and self.spacing is not None and self.physical_size is not None:
            self.size = tuple(int(round(p / s)) for p, s in zip(self.physical_size, self.spacing))  # type: ignore[assignment]


@dataclass
class Case:
    """One study: the scans (first one defines the reference grid) and the lesion annotation."""

    scans: list[Volume]
    annotation: Volume
    settings: PreprocessingSettings
    case_id: str = "case"
    scan_names: list[str] = field(default_factory=lambda: ["t2w", "adc", "hbv"])

    def write(self) -> tuple[list[Path], Path]:
        """``<scans_write_dir>/<case>_<modality index:04d>.pt`` per scan and ``<annotation_write_dir>/<case>.pt`` —
        the nnU-Net raw-data naming the reference writes (``preprocessing.py:142-173``) with tensors instead of NIfTI."""
        self.settings.scans_write_dir.mkdir(parents=True, exist_ok=True)
        self.settings.annotation_write_dir.mkdir(parents=True, exist_ok=True)
        scan_paths = []
        for index, scan in enumerate(self.scans):
            path = self.settings.scans_write_dir / f"{self.case_id}_{index:04d}.pt"
            torch.save({"data": scan.data.cpu(), "spacing": scan.spacing, "origin": scan.origin}, path)
            scan_paths.append(path)
        annotation_path = self.settings.annotation_write_dir / f"{self.case_id}.pt"
        torch.save({"data": self.annotation.data.cpu(), "spacing": self.annotation.spacing, "origin": self.annotation.origin}, annotation_path)
        return scan_paths, annotation_path

    @classmethod
    def read(cls, scan_paths: Sequence[Path], annotation_path: Path, settings: PreprocessingSettings, device: torch.device | str = "cpu") -> Case:
        def load(path: Path) -> Volume:
            blob = torch.load(path, weights_only=False)
            return Volume(blob["data"].to(device), tuple(blob["spacing"]), tuple(blob["origin"]))

        return cls([load(p) for p in scan_paths], load(annotation_path), settings, case_id=Path(annotation_path).stem)


class PreprocessingError(Exception):
    pass


def resample(volume: Volume, size: tuple[int, int, int], spacing: Triple, origin: Triple, nearest: bool = False) -> Volume:
    """Sample ``volume`` on the grid (``size``, ``spacing``, ``origin``).  Output voxel ``i`` sits at physical
    ``origin + i * spacing``; the value there is interpolated from the source (trilinear, or nearest for label maps);
    positions outside the source read 0."""
    source = volume.data
    dtype = source.dtype
    coords = []
    for axis in range(3):
        physical = origin[axis] + torch.arange(size[axis], device=source.device, dtype=torch.float32) * spacing[axis]
        index = (physical - volume.origin[axis]) / volume.spacing[axis]            # fractional source index
        n = source.shape[axis]
        coords.append(2.0 * index / max(n - 1, 1) - 1.0 if n > 1 else torch.zeros_like(index))  # align_corners=True
    dd, hh, ww = torch.meshgrid(coords[0], coords[1], coords[2], indexing="ij")
    grid = torch.stack((ww, hh, dd), dim=-1).unsqueeze(0)                           # grid_sample wants (x, y, z)
    out = F.grid_sample(source.float()[None, None], grid, mode="nearest" if nearest else "bilinear", padding_mode="zeros", align_corners=True)[0, 0]
    if not dtype.is_floating_point:
        out = out.round().to(dtype)
    return Volume(out, spacing, origin)


class PreprocessingTransform(ABC):
    @abstractmethod
    def __call__(self, case: Case) -> Case:
        raise NotImplementedError


class ResampleToFirstScan(PreprocessingTransform):
    """Bring every other sequence (ADC, HBV are acquired on a coarser grid than T2W) and the annotation onto the first
    scan's grid (parity: ``preprocessing.py:200-223``)."""

    def __call__(self, case: Case) -> Case:
        first = case.scans[0]
        scans = [first] + [resample(s, first.size, first.spacing, first.origin) for s in case.scans[1:]]
        return replace(case, scans=scans, annotation=resample(case.annotation, first.size, first.spacing, first.origin, nearest=True))


class ResampleSpacing(PreprocessingTransform):
    """Resample to ``settings.spacing`` keeping the physical extent (parity: ``preprocessing.py:226-244``)."""

    def __call__(self, case: Case) -> Case:
        target = case.settings.spacing
        if target is None:
            return case

        def to_spacing(volume: Volume, nearest: bool) -> Volume:
            size = tuple(max(int(round(n * s / t)), 1) for n, s, t in zip(volume.size, volume.spacing, target))
            return resample(volume, size, target, volume.origin, nearest)  # type: ignore[arg-type]

        return replace(case, scans=[to_spacing(s, False) for s in case.scans], annotation=to_spacing(case.annotation, True))


def centre_crop_or_pad(volume: Volume, size: tuple[int, int, int]) -> Volume:
    data, origin = volume.data, list(volume.origin)
    for axis in range(3):
        have, want = data.shape[axis], size[axis]
        if have > want:
            start = (have - want) // 2
            data = data.narrow(axis, start, want)
            origin[axis] += start * volume.spacing[axis]
        elif have < want:
            before = (want - have) // 2
            pad = [0, 0, 0, 0, 0, 0]
            pad[2 * (2 - axis)], pad[2 * (2 - axis) + 1] = before, want - have - before
            data = F.pad(data, pad)
            origin[axis] -= before * volume.spacing[axis]
    return Volume(data.contiguous(), volume.spacing, (origin[0], origin[1], origin[2]))


class CentreCropAndOrPad(PreprocessingTransform):
    """Centre crop and / or zero pad to ``settings.size`` (parity: ``preprocessing.py:247-269``)."""

    def __call__(self, case: Case) -> Case:
        size = case.settings.size
        if size is None:
            return case
        return replace(case, scans=[centre_crop_or_pad(s, size) for s in case.scans], annotation=centre_crop_or_pad(case.annotation, size))


class AlignOriginAndDirection(PreprocessingTransform):
    """After resampling, sequences can differ in origin by rounding only; snap them (and the annotation) to the first
    scan's origin, and refuse cases whose grids genuinely disagree (parity: ``preprocessing.py:272-308``)."""

    def __init__(self, tolerance_mm: float = 1e-2) -> None:
        self.tolerance_mm = tolerance_mm

    def __call__(self, case: Case) -> Case:
        first = case.scans[0]
        for volume in [*case.scans[1:], case.annotation]:
            if volume.size != first.size:
                raise PreprocessingError(f"{case.case_id}: sizes differ after resampling ({volume.size} vs {first.size})")
            if any(abs(a - b) > self.tolerance_mm + 0.5 * s for a, b, s in zip(volume.origin, first.origin, first.spacing)):
                raise PreprocessingError(f"{case.case_id}: origins differ by more than half a voxel")
        snap = lambda v: Volume(v.data, first.spacing, first.origin)  # noqa: E731
        return replace(case, scans=[first] + [snap(s) for s in case.scans[1:]], annotation=snap(case.annotation))


class BinarizeAnnotation(PreprocessingTransform):
    """ISUP-graded lesion labels (1…5) -> {0, 1} (parity: ``preprocessing.py:311-333``)."""

    def __call__(self, case: Case) -> Case:
        data = case.annotation.data
        return replace(case, annotation=Volume((data >= 1).to(data.dtype), case.annotation.spacing, case.annotation.origin))


class ZScoreNormalise(PreprocessingTransform):
    """Per-scan intensity normalisation over the non-zero voxels (what nnU-Net's MRI normalisation scheme does)."""

    def __call__(self, case: Case) -> Case:
        def normalise(volume: Volume) -> Volume:
            data = volume.data.float()
            mask = data != 0
            if not bool(mask.any()):
                return Volume(data, volume.spacing, volume.origin)
            mean, std = data[mask].mean(), data[mask].std().clamp_min(1e-6)
            return Volume(torch.where(mask, (data - mean) / std, data), volume.spacing, volume.origin)

        return replace(case, scans=[normalise(s) for s in case.scans])


def default_transforms() -> list[PreprocessingTransform]:
    return [ResampleToFirstScan(), AlignOriginAndDirection(), ResampleSpacing(), CentreCropAndOrPad(), BinarizeAnnotation()]


def apply_transform(case: Case, transforms: Sequence[PreprocessingTransform]) -> tuple[list[Path], Path]:
    for transform in transforms:
        case = transform(case)
    return case.write()


def preprocess(cases: Sequence[Case], transforms: Sequence[PreprocessingTransform] | None = None) -> list[tuple[list[Path], Path]]:
    """Run the pipeline over a study list; cases that fail a consistency check are reported and skipped, as in the
    reference (``preprocessing.py:361-384``)."""
    from logging import WARNING

    from fl4health_b200.common.logger import log

    done = []
    for case in cases:
        try:
            done.append(apply_transform(case, transforms if transforms is not None else default_transforms()))
        except PreprocessingError as error:
            log(WARNING, f"skipping {case.case_id}: {error}")
    return done
