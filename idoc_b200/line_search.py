"""Backtracking line search of iLQR / biLQR — mirror of idoc/line_search.py:7-42.

The reference returns a closure that is called inside the iLQR loop; here the loop runs on the device
(csrc/ilqr.cuh: ilqr_update_kernel + ilqr_ls_kernel), so `make_line_search` returns the parameters the kernels
need.  Acceptance rule and the accept-the-last-trial-on-failure behaviour are the reference's."""
from typing import NamedTuple


class LineSearch(NamedTuple):
    lower: float = 1e-3
    upper: float = 2.0
    rho: float = 0.5
    maxiter: int = 8
    verbose: bool = False


def make_line_search(*, lower: float = 1e-3, upper: float = 2.0, rho: float = 0.5, maxiter: int = 8,
                     verbose: bool = False) -> LineSearch:
    """Makes the line search used by ilqr and bilqr (idoc/line_search.py:7)."""
    return LineSearch(lower, upper, rho, maxiter, verbose)
