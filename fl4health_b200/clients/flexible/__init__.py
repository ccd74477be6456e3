from fl4health_b200.clients.flexible.base import FlexibleClient

__all__ = ["FlexibleClient"]
