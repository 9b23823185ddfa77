from ...modules import GraphLevelOutput, GraphLevelOutputAtoms, LSTMCellHidden, Set2Vec  # noqa: F401

__all__ = ["GraphLevelOutput", "LSTMCellHidden", "Set2Vec"]
