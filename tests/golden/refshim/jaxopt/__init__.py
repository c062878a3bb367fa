"""jaxopt stand-in: forward semantics only (custom_root = identity on the solver)."""
from . import implicit_diff, linear_solve, loop  # noqa: F401
