# Mirrors reference models/__init__.py:1 (`from models.fe import *`).
from .fe import *  # noqa: F401,F403
