from fl4health_b200.clients.mkmmd_clients.ditto_mkmmd_client import DittoMkMmdClient
from fl4health_b200.clients.mkmmd_clients.mr_mtl_mkmmd_client import MrMtlMkMmdClient

__all__ = ["DittoMkMmdClient", "MrMtlMkMmdClient"]
