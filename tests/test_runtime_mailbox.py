"""Native shared-memory mailbox (runtime/csrc/shm_mailbox.cpp): ordering, parity buffers, variable record lengths and
the timeout path, exercised with real processes."""

import multiprocessing as mp
import os

import numpy as np
import pytest

from fl4health_b200.runtime.mailbox import ShmMailbox, load_runtime

pytestmark = pytest.mark.skipif(load_runtime() is None, reason="no C++ compiler to build the host runtime")


def _worker(name: str, world: int, rank: int, iterations: int, queue) -> None:  # noqa: ANN001
    box = ShmMailbox(name, world, rank, capacity=64)
    bad = 0
    rng = np.random.default_rng(rank)
    for it in range(1, iterations + 1):
        if rng.random() < 0.05:
            os.sched_yield()
        n = 1 + (it + rank) % 40
        records = box.all_gather(np.full(n, it * 1000.0 + rank))
        for r, rec in enumerate(records):
            if len(rec) != 1 + (it + r) % 40 or not np.all(rec == it * 1000.0 + r):
                bad += 1
    box.close()
    queue.put((rank, bad))


def test_mailbox_all_gather_across_processes() -> None:
    world, iterations = 3, 2000
    name = f"/fl4h_test_{os.getpid()}"
    owner = ShmMailbox(name, world, 0, capacity=64, create=True)
    ctx = mp.get_context("fork")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(name, world, r, iterations, queue)) for r in range(1, world)]
    for p in procs:
        p.start()
    bad = 0
    try:
        for it in range(1, iterations + 1):
            n = 1 + it % 40
            records = owner.all_gather(np.full(n, it * 1000.0), timeout=60.0)
            for r, rec in enumerate(records):
                if len(rec) != 1 + (it + r) % 40 or not np.all(rec == it * 1000.0 + r):
                    bad += 1
    finally:
        owner.unlink()
    results = [queue.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert bad == 0 and all(b == 0 for _, b in results)
    owner.close()


def test_mailbox_timeout_and_capacity() -> None:
    name = f"/fl4h_test_t_{os.getpid()}"
    box = ShmMailbox(name, 2, 0, capacity=8, create=True)
    box.unlink()
    with pytest.raises(ValueError):
        box.all_gather(np.zeros(9))
    with pytest.raises(TimeoutError):  # rank 1 never posts
        box.all_gather(np.zeros(2), timeout=0.2)
    box.close()
    with pytest.raises(RuntimeError):
        ShmMailbox(name, 2, 1, capacity=8)  # the name is gone
