from ...modules import AtomAutoEncoder  # noqa: F401
