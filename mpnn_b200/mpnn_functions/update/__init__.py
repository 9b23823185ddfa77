from ...modules import GRUCell, GRUUpdate  # noqa: F401

__all__ = ["GRUUpdate"]
