from .._transport import start_client  # noqa: F401
from .numpy_client import Client, NumPyClient  # noqa: F401


def start_numpy_client(*, server_address: str, client: NumPyClient, **kwargs) -> None:  # noqa: ANN003
    start_client(server_address=server_address, client=client.to_client(), **kwargs)
