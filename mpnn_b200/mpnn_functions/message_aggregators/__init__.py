from ...modules import AdjMsgAgg, AttMsgAgg, WAdjMsgAgg  # noqa: F401

__all__ = ["AdjMsgAgg", "WAdjMsgAgg", "AttMsgAgg"]
