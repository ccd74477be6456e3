"""Batched, pinned, prefetching loader for in-memory tensor datasets.

The reference feeds clients with ``DataLoader(num_workers=0, pin_memory=False)`` over a dataset whose
``__getitem__`` runs a Python transform per sample, then a blocking ``.to(device)`` per batch
(``fl4health/utils/load_data.py:244-245``, ``utils/client.py:61-64``).  CIFAR-10 is 150 MB; a B200 has 180 GB of HBM.
``BatchedTensorLoader`` therefore:

* gathers a whole batch with one ``index_select`` (no per-sample Python);
* ``placement="pinned"``: keeps the dataset in page-locked host memory and stages each batch into a ring of pinned
  buffers so the engine's H2D copy is asynchronous and overlaps the previous step;
* ``placement="device"``: keeps the whole dataset resident in HBM — batches never touch the host;
* exposes ``dataset`` / ``batch_size`` / ``__len__`` like a ``DataLoader`` so ``BasicClient`` code is unchanged.
"""

from __future__ import annotations

from collections.abc import Iterator

import torch

from fl4health_b200.utils.dataset import TensorDataset


class BatchedTensorLoader:
    def __init__(
        self,
        dataset: TensorDataset,
        batch_size: int,
        shuffle: bool = False,
        drop_last: bool = False,
        placement: str = "host",
        device: torch.device | str | None = None,
        generator: torch.Generator | None = None,
        ring: int = 4,
    ) -> None:
        assert placement in ("host", "pinned", "device")
        self.dataset = dataset
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.placement = placement
        self.device = torch.device(device) if device is not None else None
        self.generator = generator
        self._ring_size = ring
        self._ring: list[tuple[torch.Tensor, torch.Tensor]] = []
        self._ring_pos = 0
        if placement == "device":
            assert self.device is not None, "device placement needs a device"
            dataset.data = dataset.data.to(self.device)
            if dataset.targets is not None:
                dataset.targets = dataset.targets.to(self.device)
        elif placement == "pinned" and torch.cuda.is_available():
            dataset.data = dataset.data.pin_memory()
            if dataset.targets is not None:
                dataset.targets = dataset.targets.pin_memory()

    def __len__(self) -> int:
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _stage(self, data: torch.Tensor, target: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Copy a freshly gathered host batch into the next pinned ring slot (keeps H2D copies async)."""
        if self.placement != "pinned" or not torch.cuda.is_available():
            return data, target
        if not self._ring or self._ring[0][0].shape != data.shape:
            self._ring = [
                (torch.empty_like(data).pin_memory(), torch.empty_like(target).pin_memory())
                for _ in range(self._ring_size)
            ]
            self._ring_pos = 0
        if data.shape != self._ring[0][0].shape:
            return data.pin_memory(), target.pin_memory()
        slot_d, slot_t = self._ring[self._ring_pos]
        self._ring_pos = (self._ring_pos + 1) % self._ring_size
        slot_d.copy_(data)
        slot_t.copy_(target)
        return slot_d, slot_t

    def __iter__(self) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
        n = len(self.dataset)
        index_device = self.dataset.data.device if self.placement == "device" else torch.device("cpu")
        if self.shuffle:
            order = torch.randperm(n, generator=self.generator).to(index_device)
        else:
            order = torch.arange(n, device=index_device)
        for start in range(0, n, self.batch_size):
            idx = order[start : start + self.batch_size]
            if idx.numel() < self.batch_size and self.drop_last:
                return
            data, target = self.dataset.get_batch(idx)
            if idx.numel() == self.batch_size:
                data, target = self._stage(data, target)
            yield data, target
