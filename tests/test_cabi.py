"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol `include/d3b200.h` declares, and the host layer refuses to run
without a GPU instead of falling back."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tad_dftd3_b200 import build, _lib

    build.build()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "d3b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d3b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from tad_dftd3_b200 import _lib

    names = declared_symbols()
    assert names, "no declarations found in include/d3b200.h"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in d3b200.h but not exported"
    # the ctypes table binds exactly the declared set
    assert sorted(_lib.exported_symbols()) == names


def test_abi_version_and_limits(lib):
    assert lib.d3b200_abi_version() == 1
    assert lib.d3b200_batch_max_atoms_f64(0) >= 512
    assert lib.d3b200_batch_max_atoms_f64(1) >= 256
    assert lib.d3b200_batch_max_atoms_f32(0) >= lib.d3b200_batch_max_atoms_f64(0)


def test_argument_errors_without_gpu(lib):
    """Argument validation happens before any CUDA call."""
    from tad_dftd3_b200 import _lib
    import ctypes as C

    par, model = _lib.Params(), _lib.Model()
    rc = lib.d3b200_batch_energy_f64(1, 4, None, None, C.byref(model), C.byref(par), None, None, 0, None)
    assert rc == 1 and b"null" in lib.d3b200_last_error()


def test_no_cpu_fallback():
    import tad_dftd3_b200 as d3

    numbers = torch.tensor([1, 1])
    positions = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.4]])
    with pytest.raises(RuntimeError, match="CUDA"):
        d3.dftd3(numbers, positions, {"a1": torch.tensor(0.4)})


def test_argument_validation_matches_reference():
    """ValueError for Z > 103 and shape mismatch (reference test_dftd3.py:31-50)."""
    import tad_dftd3_b200 as d3

    positions = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    with pytest.raises(ValueError):
        d3.dftd3(torch.tensor([1, 105]), positions, {})
    with pytest.raises(ValueError):
        d3.dftd3(torch.tensor([1]), positions, {})
    with pytest.raises(NotImplementedError):
        d3.dftd3(torch.tensor([1, 1]), positions, {}, counting_function=lambda r, r0: r)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from tad_dftd3_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.D3LibraryError, match="no CPU"):
        _lib.lib()
