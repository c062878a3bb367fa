"""Solution / solver types — same names and field order as the reference, idoc/typs.py:6-16."""
from dataclasses import dataclass
from typing import Callable, NamedTuple, Optional

import numpy as np


class State(NamedTuple):
    X: np.ndarray
    U: np.ndarray
    Nu: np.ndarray


@dataclass
class Solver:
    direct: Callable
    implicit: Callable
    kkt: Callable
    # Extension (the reference obtains this through jax.vjp of `implicit`): vjp(params) -> (State, pullback),
    # pullback(State cotangent) -> Params cotangent, computed by the exact adjoint-LQR kernels.
    vjp: Optional[Callable] = None
