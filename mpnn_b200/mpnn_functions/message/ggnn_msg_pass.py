from ...modules import GGNNMsgPass  # noqa: F401
