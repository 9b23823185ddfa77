"""Drop-in for the reference's `mpnn_functions/encoders` package: same classes, layers and
state_dict keys; the `.encoder` halves evaluate in row space on categorical bond tensors (mpnn_b200.modules)."""
from .auto_encoder import Autoencoder  # noqa: F401
