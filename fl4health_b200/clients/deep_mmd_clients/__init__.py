from fl4health_b200.clients.deep_mmd_clients.ditto_deep_mmd_client import DittoDeepMmdClient
from fl4health_b200.clients.deep_mmd_clients.mr_mtl_deep_mmd_client import MrMtlDeepMmdClient

__all__ = ["DittoDeepMmdClient", "MrMtlDeepMmdClient"]
