"""mapanything_b200: B200-native (sm_100a) implementation of MapAnything's feed-forward inference hot path."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
