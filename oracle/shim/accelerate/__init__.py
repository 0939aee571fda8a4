"""ORACLE ONLY: the two names magicdrive/misc/common.py:5-8 imports from accelerate."""
from . import state, utils  # noqa: F401
