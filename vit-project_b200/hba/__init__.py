"""hba — host side of the B200 (sm_100a) CLIP-HBA-Behavior hot path.

``import hba`` requires the compiled C-ABI library (hba/libhba.so, built by
``__graft_entry__.build()``); there is no CPU fallback.
"""
from . import _lib

_lib.load()  # fail loudly at import time when the CUDA extension is missing

from . import ops  # noqa: E402
from .engine import get_engine, get_precision, set_precision  # noqa: E402
from .dora import DoRALayer  # noqa: E402

__all__ = ["ops", "get_engine", "get_precision", "set_precision", "DoRALayer"]
