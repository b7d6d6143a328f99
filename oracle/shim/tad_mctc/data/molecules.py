"""
`tad_mctc.data.molecules` -- the molecule collection of tad-mctc is NOT
available offline.  `mols` is empty; `oracle/make_golden.py` attaches the few
geometries that exist in the reference tree itself (examples/, README).
"""
mols: dict = {}


def merge_nested_dicts(a: dict, b: dict) -> dict:
    out = {k: dict(v) for k, v in a.items()}
    for k, v in b.items():
        out.setdefault(k, {}).update(v)
    return out
