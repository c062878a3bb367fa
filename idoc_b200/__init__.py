"""idoc_b200 — B200-native (sm_100a) batched LQR hot path behind the API of tachukao/idoc.

Modules mirror the reference package (idoc/__init__.py:1-15): lqr, typs, utils .  Importing requires the built CUDA library
(idoc_b200/libidoc_b200.so); there is no CPU fallback.
"""
from . import _lib, typs, utils, lqr, tilqr, ilqr, blqr, bilqr, line_search, host, dist  # noqa: F401
from .line_search import make_line_search  # noqa: F401

__all__ = ["lqr", "tilqr", "ilqr", "blqr", "bilqr", "typs", "utils", "line_search", "make_line_search", "host", "dist"]
