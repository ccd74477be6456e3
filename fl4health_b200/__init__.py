"""fl4health_b200 — a Blackwell-native federated-learning engine.

Same capability surface as VectorInstitute/FL4Health (BasicClient hooks, FlServer family, strategies,
exchangers, checkpoint layout) but built B200-first: one process per GPU, flat symmetric parameter arenas,
hand-written sm_100a kernels for the exchange / aggregate / optimizer / GEMM hot paths, CUDA graphs for the
local training step, NCCL only for plumbing.
"""

__version__ = "0.1.0"
