"""genomeworks.cudapoa -> genomeworks_b200.cudapoa"""
from genomeworks_b200.cudapoa import *  # noqa: F401,F403
from genomeworks_b200 import cudapoa as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
