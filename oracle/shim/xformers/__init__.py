"""ORACLE ONLY: stand-in for xformers (un-vendored, version unpinned: reference README.md:70)."""
from . import ops  # noqa: F401
