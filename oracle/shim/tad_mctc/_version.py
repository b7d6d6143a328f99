"""torch version as a tuple (tad-mctc ``_version.__tversion__``)."""
import re

import torch


def _parse(v: str):
    m = re.match(r"(\d+)\.(\d+)\.(\d+)", v)
    assert m is not None
    return tuple(int(x) for x in m.groups())


__tversion__ = _parse(torch.__version__)
