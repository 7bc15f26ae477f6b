"""kbot-joystick_b200 -- B200-native (sm_100a) rollout control step of the kbot-joystick task.

Host side is thin: `engine.KbotStep` binds the C-ABI of libkbotstep.so (include/kbotstep.h) with ctypes and
`task.HumanoidWalkingTask` mirrors the reference's ksim Task plugin surface (train.py:1058-1756) on top of it.
Importing the package does not need a GPU; any compute call does (no CPU fallback).
"""

from . import _lib, spec  # noqa: F401
from ._lib import LIB_PATH  # noqa: F401

__all__ = ["_lib", "spec", "LIB_PATH"]
