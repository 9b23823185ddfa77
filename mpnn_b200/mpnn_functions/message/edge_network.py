from ...modules import EdgeNetwork  # noqa: F401
