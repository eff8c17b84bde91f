"""boxdreamer_b200 -- B200-native (sm_100a) implementation of BoxDreamer's inference hot path.

Public surface mirrors the reference's `src/models` module API:
    BoxDreamer(config).forward(data: dict) -> dict      (src/models/BoxDreamerModel.py:21,112-191)
The arithmetic lives in `libboxdreamer_b200.so` (C ABI: include/boxdreamer_b200.h), built in-tree by
`python -m boxdreamer_b200.build`.
"""
from .model import BETR, BoxDreamer, DinoV2Wrapper, Engine  # noqa: F401
from . import _lib, synth  # noqa: F401

__version__ = "0.1.0"
