from ...modules import AttEdgeNetwork  # noqa: F401
