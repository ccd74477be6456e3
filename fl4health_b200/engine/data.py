"""Batched, pinned, prefetching loader for in-memory tensor datasets.

The reference feeds clients with ``DataLoader(num_workers=0, pin_memory=False)`` over a dataset whose
``__getitem__`` runs a Python transform per sample, then a blocking ``.to(device)`` per batch
(``fl4health/utils/load_data.py:244-245``, ``utils/client.py:61-64``).  CIFAR-10 is 150 MB; a B200 has 180 GB of HBM.
``BatchedTensorLoader`` therefore:

* gathers a whole batch with one ``index_select`` (no per-sample Python);
* ``placement="pinned"``: keeps the dataset in page-locked host memory.  Shuffled epochs are materialised ONCE per
  epoch (one permuted gather into a second pinned buffer, prepared by a background thread while the previous epoch is
  being consumed) and batches are contiguous *views* of that buffer: no per-batch gather, no staging copy, and the
  engine's H2D copy is asynchronous.  (Measured on the B200 box: a per-batch ``index_select`` of 32 CIFAR rows from
  pinned memory costs ~1 ms of host time -- more than the whole GPU training step.)
* ``placement="device"``: keeps the whole dataset resident in HBM — batches never touch the host;
* exposes ``dataset`` / ``batch_size`` / ``__len__`` like a ``DataLoader`` so ``BasicClient`` code is unchanged.
"""

from __future__ import annotations

from collections.abc import Iterator
from concurrent.futures import Future, ThreadPoolExecutor

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
        # epoch-level shuffling (host / pinned placements): two epoch buffers, the next one filled in the background
        self._epoch_bufs: list[tuple[torch.Tensor, torch.Tensor] | None] = [None, None]
        self._epoch_next: tuple[int, Future] | None = None
        self._epoch_pool: ThreadPoolExecutor | None = None
        self._epoch_done: dict[int, torch.cuda.Event] = {}  # H2D copies out of an epoch buffer have been enqueued
        self.max_epoch_buffer_bytes = 8 << 30
        # pinned placement + device: copy-stream prefetch (see _iter_device_prefetched)
        self.prefetch_to_device = True
        self.prefetch_depth = 2
        self._copy_stream: torch.cuda.Stream | None = None
        self._device_ring: list[tuple[torch.Tensor, torch.Tensor]] = []
        self._device_ring_pos = 0
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

    # ------------------------------------------------------------------------------------------------------
    def _plain(self) -> bool:
        """True when batches are exactly ``apply_transforms(data[idx], targets[idx])`` (no dataset-level override)."""
        ds = self.dataset
        return (
            type(ds).get_batch is TensorDataset.get_batch and getattr(ds, "targets", None) is not None
            and isinstance(ds.data, torch.Tensor) and isinstance(ds.targets, torch.Tensor)
        )

    def _epoch_mode(self) -> bool:
        if not (self.shuffle and self.placement in ("host", "pinned") and self._plain()):
            return False
        ds = self.dataset
        nbytes = ds.data.numel() * ds.data.element_size()
        return len(self) >= 8 and nbytes <= self.max_epoch_buffer_bytes

    def _fill_epoch(self, slot: int, perm: torch.Tensor) -> int:
        ds = self.dataset
        bufs = self._epoch_bufs[slot]
        if bufs is None or bufs[0].shape != ds.data.shape or bufs[0].dtype != ds.data.dtype:
            pin = self.placement == "pinned" and torch.cuda.is_available()
            data_buf, target_buf = torch.empty_like(ds.data), torch.empty_like(ds.targets)
            bufs = (data_buf.pin_memory(), target_buf.pin_memory()) if pin else (data_buf, target_buf)
            self._epoch_bufs[slot] = bufs
        done = self._epoch_done.pop(slot, None)
        if done is not None:
            done.synchronize()  # asynchronous H2D copies may still be reading this buffer's previous contents
        torch.index_select(ds.data, 0, perm, out=bufs[0])
        torch.index_select(ds.targets, 0, perm, out=bufs[1])
        return slot

    def _iter_epoch_buffered(self) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
        n = len(self.dataset)
        if self._epoch_pool is None:
            self._epoch_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="fl4h-epoch")
        if self._epoch_next is not None:
            slot = self._epoch_next[1].result()
        else:
            slot = self._fill_epoch(0, torch.randperm(n, generator=self.generator))
        # permutations are drawn on the caller's thread (deterministic RNG consumption); only the gather is offloaded
        upcoming = torch.randperm(n, generator=self.generator)
        self._epoch_next = (1 - slot, self._epoch_pool.submit(self._fill_epoch, 1 - slot, upcoming))
        data, targets = self._epoch_bufs[slot]  # type: ignore[misc]
        try:
            for start in range(0, n, self.batch_size):
                stop = min(start + self.batch_size, n)
                if stop - start < self.batch_size and self.drop_last:
                    return
                yield self.dataset.apply_transforms(data[start:stop], targets[start:stop])
        finally:
            if self.placement == "pinned" and torch.cuda.is_available():
                event = torch.cuda.Event()
                event.record()
                self._epoch_done[slot] = event

    def _iter_sequential(self) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
        ds, n = self.dataset, len(self.dataset)
        for start in range(0, n, self.batch_size):
            stop = min(start + self.batch_size, n)
            if stop - start < self.batch_size and self.drop_last:
                return
            yield ds.apply_transforms(ds.data[start:stop], ds.targets[start:stop])

    def __iter__(self) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
        if self.placement == "pinned" and self.device is not None and self.device.type == "cuda" and self.prefetch_to_device:
            return self._iter_device_prefetched(self._iter_host())
        return self._iter_host()

    def _iter_host(self) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
        if not self.shuffle and self._plain():
            return self._iter_sequential()  # contiguous views: zero-copy for every placement
        if self._epoch_mode():
            return self._iter_epoch_buffered()
        return self._iter_gathered()

    def _iter_device_prefetched(self, host_batches: Iterator) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
        """Pinned placement with a target device: H2D copies run on a dedicated copy stream ``prefetch_depth`` batches
        ahead of the consumer, into a small ring of device buffers; the consumer's stream only waits on the copy's event
        (already complete in steady state), so the host->device transfer never sits on the training critical path."""
        depth = self.prefetch_depth
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        copy_stream = self._copy_stream
        in_flight: list[tuple[torch.Tensor, torch.Tensor, torch.cuda.Event]] = []

        def enqueue(batch: tuple[torch.Tensor, torch.Tensor]) -> None:
            data, target = batch
            slot = self._device_ring_pos % (depth + 2)
            self._device_ring_pos += 1
            if len(self._device_ring) <= slot or self._device_ring[slot][0].shape != data.shape:
                fresh = (torch.empty(data.shape, dtype=data.dtype, device=self.device),
                         torch.empty(target.shape, dtype=target.dtype, device=self.device))
                if len(self._device_ring) <= slot:
                    self._device_ring.append(fresh)
                else:
                    self._device_ring[slot] = fresh
            dst_data, dst_target = self._device_ring[slot]
            # the slot's previous consumer ran on the current stream: the copy must not overtake it
            copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(copy_stream):
                dst_data.copy_(data, non_blocking=True)
                dst_target.copy_(target, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            in_flight.append((dst_data, dst_target, ready))

        for batch in host_batches:
            if not (isinstance(batch[0], torch.Tensor) and isinstance(batch[1], torch.Tensor)):
                yield batch  # structured batches (dict inputs) take the plain path
                continue
            enqueue(batch)
            if len(in_flight) > depth:
                data, target, ready = in_flight.pop(0)
                torch.cuda.current_stream(self.device).wait_event(ready)
                yield data, target
        while in_flight:
            data, target, ready = in_flight.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ready)
            yield data, target

    def _iter_gathered(self) -> Iterator[tuple[torch.Tensor, torch.Tensor]]:
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
