"""Stand-alone substitute for the subset of Flower (flwr 1.18) that FL4Health's stock code path touches.

Written for the benchmark's reference arm only: the image has no ``flwr`` wheel, so the unmodified reference package in
``baseline/_ref/fl4health`` is driven through this shim.  Nothing here imports ``fl4health_b200``.  The gRPC transport
is replaced by a length-prefixed pickle stream over localhost TCP (``flwr._transport``): parameters still travel as
``np.save`` byte blobs between a server process and one client process per GPU, exactly the reference's shape of work
minus protobuf / HTTP2 framing (i.e. the shim is, if anything, cheaper than the real wire).
"""

from . import client, common, server  # noqa: F401

__version__ = "1.18.0+shim"
