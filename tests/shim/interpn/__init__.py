"""`import interpn` -> the B200 backend (drop-in alias used to run the reference's own Python tests unmodified).

The published `interpn` wheel is built with the crate's `fma` feature (/root/reference/pyproject.toml:72), so the alias
selects the `fma` flavour of the CUDA library unless INTERPN_B200_ARITHMETIC says otherwise; everything else is the
public surface of /root/reference/src/interpn/__init__.py:1-46 re-exported from interpn_b200.
"""

import os as _os

_os.environ.setdefault("INTERPN_B200_ARITHMETIC", "fma")

import interpn_b200 as _ib  # noqa: E402
from interpn_b200 import (  # noqa: E402,F401
    MulticubicRectilinear,
    MulticubicRegular,
    MultilinearRectilinear,
    MultilinearRegular,
    NearestRectilinear,
    NearestRegular,
    interpn,
    one_dim,
    raw,
)

__version__ = "0.8.2"  # the reference release whose surface is mirrored
__backend__ = f"interpn_b200 {_ib.__version__} ({_ib._lib.ARITHMETIC})"

__all__ = [
    "__version__",
    "interpn",
    "raw",
    "MultilinearRegular",
    "MultilinearRectilinear",
    "MulticubicRegular",
    "MulticubicRectilinear",
    "NearestRegular",
    "NearestRectilinear",
]
