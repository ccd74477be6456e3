"""CPU / NUMA placement of a rank process next to its GPU.

On a two-socket 8-GPU box the ranks of a job otherwise float over all cores: page-locked staging buffers end up on the
far socket, the epoch-gather thread of one rank competes with the main thread of another, and — because a federated
round is a max over ranks with host rendezvous points — the slowest rank sets the pace (the round-1 scaling run lost
32 % end to end at 8 GPUs with device-resident numbers unaffected).  ``bind_to_gpu`` restricts the calling process to
its GPU's NUMA node and, within the node, to an equal share of the cores among the GPUs attached to that node; call it
BEFORE allocating pinned memory or starting worker threads (first-touch page placement follows the binding).
"""

from __future__ import annotations

import os
from logging import INFO
from pathlib import Path

import torch

from fl4health_b200.common.logger import log


def _parse_cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> int | None:
    try:
        props = torch.cuda.get_device_properties(device_index)
        address = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{address}/numa_node").read_text().strip())
        return node if node >= 0 else None
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def bind_to_gpu(device_index: int) -> dict:
    """Returns what was done (for logs / bench JSON); a no-op description when the topology cannot be read."""
    if os.environ.get("FL4H_AFFINITY", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return {"bound": False, "reason": "disabled"}
    node = gpu_numa_node(device_index)
    if node is None:
        return {"bound": False, "reason": "no NUMA information for the GPU"}
    try:
        node_cpus = _parse_cpulist(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text())
    except OSError:
        return {"bound": False, "reason": f"cannot read cpulist of node {node}"}
    allowed = sorted(set(node_cpus) & os.sched_getaffinity(0))
    if not allowed:
        return {"bound": False, "reason": "GPU node has no CPU this process may use"}
    # equal share among the visible GPUs attached to the same node (hyper-thread siblings stay together when the
    # kernel enumerates them as the upper half of the list: a contiguous slice of each half)
    peers = [i for i in range(torch.cuda.device_count()) if gpu_numa_node(i) == node]
    share = allowed
    if len(peers) > 1 and device_index in peers and len(allowed) >= 2 * len(peers):
        position, per = peers.index(device_index), len(allowed) // len(peers)
        share = allowed[position * per : (position + 1) * per]
    os.sched_setaffinity(0, share)
    torch.set_num_threads(max(1, min(len(share) // 2, 8)))  # gathers / pinned copies may use a few cores of the share
    log(INFO, f"rank process bound to NUMA node {node}: {len(share)} CPUs ({share[0]}..{share[-1]}) for cuda:{device_index}")
    return {"bound": True, "numa_node": node, "cpus": len(share), "first_cpu": share[0], "last_cpu": share[-1]}
