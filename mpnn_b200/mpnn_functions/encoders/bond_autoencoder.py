from ...modules import BondAutoEncoder  # noqa: F401
