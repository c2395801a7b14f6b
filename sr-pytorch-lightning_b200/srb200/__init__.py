"""srb200 — host layer of the B200-native SR conv hot path (ctypes -> libsrb200.so)."""
from . import lib  # noqa: F401

__all__ = ["lib", "ops", "functional"]
