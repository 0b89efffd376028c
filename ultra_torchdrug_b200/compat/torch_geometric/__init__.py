"""`torch_geometric` stand-in: only `utils.remove_self_loops` is used (reference util.py:19)."""
from . import utils  # noqa: F401
