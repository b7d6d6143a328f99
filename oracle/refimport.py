"""
Import the UNMODIFIED reference (`/root/reference/src/tad_dftd3`, or the copy
pip-installed under `baseline/_ref`) on top of the tad-mctc stand-in.

TEST INFRASTRUCTURE ONLY: used by `oracle/make_golden.py`, by tests that run
in the build container, and by `bench.py --impl reference`.  Never imported by
`tad_dftd3_b200`.
"""
from __future__ import annotations

import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def reference_available() -> str | None:
    """Path that provides `tad_dftd3`, or None."""
    for p in (
        os.path.join(_ROOT, "baseline", "_ref"),
        "/root/reference/src",
    ):
        if os.path.isfile(os.path.join(p, "tad_dftd3", "disp.py")):
            return p
    return None


def import_reference():
    """Return the reference `tad_dftd3` module (raises ImportError if absent)."""
    path = reference_available()
    if path is None:
        raise ImportError("reference tad_dftd3 not found (baseline/_ref, /root/reference)")
    shim = os.path.join(_HERE, "shim")
    try:
        import tad_mctc  # noqa: F401  (a real install wins over the stand-in)
    except ImportError:
        if shim not in sys.path:
            sys.path.insert(0, shim)
    if path not in sys.path:
        sys.path.append(path)
    import tad_dftd3

    return tad_dftd3
