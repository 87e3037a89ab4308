"""Host-side runtime of the B200-native scoring path: ctypes binding + engine."""
from . import native  # noqa: F401
from .engine import Engine, engine_for, invalidate  # noqa: F401
