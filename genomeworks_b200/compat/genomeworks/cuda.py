"""genomeworks.cuda -> genomeworks_b200.cuda"""
from genomeworks_b200.cuda import *  # noqa: F401,F403
from genomeworks_b200 import cuda as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
