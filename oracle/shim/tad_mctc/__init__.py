"""
Stand-in for the un-vendored dependency ``tad-mctc==0.5.3`` (pinned in the
reference's setup.cfg:37).  TEST INFRASTRUCTURE ONLY.

The reference (`/root/reference/src/tad_dftd3`) imports about a dozen symbols
from tad-mctc; that package is not installed in this image and cannot be
fetched (no network).  This package restates exactly those symbols from the
published behaviour of tad-mctc 0.5.3 so that the UNMODIFIED reference source
can be imported and run as the oracle (see ``oracle/refimport.py``).

Nothing under ``tad_dftd3_b200/`` imports this package.
"""
from . import autograd, batch, convert, data, math, ncoord, storch, tools, typing
from ._version import __tversion__

__version__ = "0.5.3+standin"
