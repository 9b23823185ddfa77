from ...modules import Autoencoder  # noqa: F401
