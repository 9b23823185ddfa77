"""Drop-in for the reference's `mpnn_functions` package (same exported names, mpnn_functions/__init__.py:1-4)."""
from .message import *  # noqa: F401,F403
from .update import *  # noqa: F401,F403
from .readout import *  # noqa: F401,F403
from .message_aggregators import *  # noqa: F401,F403
