"""Hand-written sm_100a kernels (``csrc/``) and their Python entry points.

* ``flat``  — streaming kernels over a rank's flat parameter arena: fused SGD/AdamW (+FedProx, +SCAFFOLD, +bf16
  shadow), K-way weighted aggregation with FedOpt/SCAFFOLD epilogues, broadcast unpack, drift norms, DP clip/noise.
* ``p2p``   — fused collective kernels over peer-mapped (NVLink/NVSwitch) symmetric memory.
* ``gemm``  — tcgen05/TMEM/TMA GEMM with fused bias/activation epilogues.
Every entry point has a PyTorch reference implementation used on CPU and as the numerics oracle in tests.
"""

from fl4health_b200.ops._lib import available, launch_count, load, reset_launch_count

__all__ = ["available", "launch_count", "load", "reset_launch_count"]
