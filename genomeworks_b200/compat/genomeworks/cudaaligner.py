"""genomeworks.cudaaligner -> genomeworks_b200.cudaaligner"""
from genomeworks_b200.cudaaligner import *  # noqa: F401,F403
from genomeworks_b200 import cudaaligner as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
