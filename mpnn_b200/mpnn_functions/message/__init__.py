from ...modules import AttEdgeNetwork, BiLiniearEdgeNetwork, EdgeNetwork  # noqa: F401

__all__ = ["AttEdgeNetwork", "EdgeNetwork", "BiLiniearEdgeNetwork"]
