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
    assert lib.d3b200_abi_version() == 2
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


def test_header_is_plain_c(tmp_path):
    """include/d3b200.h must be consumable from C (the drop-in boundary is a C ABI)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "d3b200.h"\nint main(void) { d3b200_params p; (void)p; return D3B200_OK; }\n')
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_stage_wrappers_validate_like_the_reference():
    """Shape errors of the stage functions (reference disp.py:212-219, tad-mctc cn_d3) before any launch."""
    import tad_dftd3_b200 as d3

    numbers = torch.tensor([1, 1, 8])
    positions = torch.zeros(3, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        d3.ncoord.cn_d3(numbers, positions)
    with pytest.raises(RuntimeError, match="CUDA"):
        d3.disp.dispersion(numbers, positions, {}, torch.zeros(3, 3), r4r2=torch.ones(3))


def test_reference_object_contract():
    """Mirror of the reference's Reference holder checks (reference.py:231-245, test_reference.py)."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import REFCN

    cn, c6 = REFCN(torch.double), syn.synthetic_c6_table(torch.double)
    ref = d3.Reference(cn=cn, c6=c6)
    assert ref.dtype == torch.double and ref.device == cn.device
    assert "n_element=104" in str(ref) and "n_reference=7" in repr(ref)
    assert ref.type(torch.double) is ref and ref.type(torch.float32).dtype == torch.float32
    with pytest.raises(RuntimeError, match="dtype"):
        d3.Reference(cn=cn.float(), c6=c6)
    with pytest.raises(RuntimeError, match="size mismatch"):
        d3.Reference(cn=cn[:50], c6=c6)
    with pytest.raises(AttributeError):
        ref.device = "cpu"


def test_data_tables():
    """R4R2 formula (data/r4r2.py:80-85), radii validated against the pinned CN elements, refcn layout."""
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN, VDW_PAIRWISE, set_vdw_pairwise

    q = R4R2(torch.double)
    assert q.shape == (119,) and q[0] == 0 and abs(float(q[1]) - (0.5 * 8.0589 * 1.0) ** 0.5) < 1e-12
    rc = COV_D3(torch.double)
    assert abs(float(rc[1]) - 4.0 / 3.0 * 0.32 / 0.529177210903) < 1e-12
    refcn = REFCN(torch.double)
    assert refcn.shape == (104, 7) and (refcn[0] == -1).all() and float(refcn[1, 0]) == 0.9118
    with pytest.raises(FileNotFoundError):
        VDW_PAIRWISE(torch.double)
    set_vdw_pairwise(torch.ones(104, 104))
    try:
        assert VDW_PAIRWISE(torch.double).shape == (104, 104)
    finally:
        set_vdw_pairwise(None)


def test_exception_and_typing_modules_mirror_the_reference():
    """tad_dftd3.exception.DFTD3Error (reference test_exception.py) and the typing aliases exist under the same names."""
    import tad_dftd3_b200 as d3

    with pytest.raises(d3.exception.DFTD3Error):
        raise d3.exception.DFTD3Error
    assert issubclass(d3.exception.DFTD3Error, Exception)
    for name in ("DD", "CountingFunction", "DampingFunction", "WeightingFunction", "Molecule", "Size", "Tensor",
                 "TensorOrTensors", "get_default_device", "get_default_dtype"):
        assert hasattr(d3.typing, name), name
    assert d3.typing.get_default_dtype() == torch.get_default_dtype()


def test_batch_order_hint_is_a_size_sorted_permutation():
    """`ops.batch_order` (host logic of d3b200_model.batch_order): a permutation, largest structures first, stable."""
    from tad_dftd3_b200 import ops

    gen = torch.Generator().manual_seed(3)
    z = (torch.rand(50, 9, generator=gen) < 0.6).long() * 6
    order = ops.batch_order(z)
    assert order.dtype == torch.int32 and sorted(order.tolist()) == list(range(50))
    counts = (z != 0).sum(-1)[order.long()]
    assert bool((counts[:-1] >= counts[1:]).all())
    same = counts[:-1] == counts[1:]
    assert bool((order[:-1][same] < order[1:][same]).all())      # stable within equal sizes


def test_model_struct_layout_matches_header():
    """The ctypes mirror of d3b200_model_* has one member per member of the C struct, in order."""
    from tad_dftd3_b200 import _lib

    text = open(os.path.join(ROOT, "include", "d3b200.h")).read()
    body = re.search(r"typedef struct d3b200_model_f64 \{(.*?)\} d3b200_model_f64;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+);", body)
    assert names == [f[0] for f in _lib.Model._fields_]
    body = re.search(r"typedef struct d3b200_system_f64 \{(.*?)\} d3b200_system_f64;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n for decl in re.findall(r"([^;]+);", body) for n in re.findall(r"(\w+)\s*(?:,|$)", decl.strip())]
    assert names == [f[0] for f in _lib.System._fields_]
    body = re.search(r"typedef struct d3b200_params \{(.*?)\} d3b200_params;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n for decl in re.findall(r"(?:double|int32_t) ([^;]+);", body) for n in re.split(r"\s*,\s*", decl.strip())]
    assert names == [f[0] for f in _lib.Params._fields_]
    import ctypes as C

    assert C.sizeof(_lib.Params) == 10 * 8 + 2 * 4


def test_staged_host_entry_chunk_schedule(lib):
    """d3b200_batch_host_chunks: boundaries start at 0, end at nbatch, strictly increase, at most 64 chunks,
    quarter / half chunks at both ends of a long batch."""
    import ctypes as C
    import random

    from tad_dftd3_b200 import HostPipeline

    rng = random.Random(0)
    for _ in range(3000):
        n = rng.randint(1, 200000)
        chunk = rng.choice([0, 1, 2, 3, 5, 16, 100, 592, 5000])
        b = (C.c_int * 65)()
        k = lib.d3b200_batch_host_chunks(n, chunk, b)
        bounds = list(b[: k + 1])
        assert 1 <= k <= 64 and bounds[0] == 0 and bounds[-1] == n
        assert all(x < y for x, y in zip(bounds, bounds[1:]))
        assert HostPipeline.chunks(n, chunk) == k
    b = (C.c_int * 65)()
    k = lib.d3b200_batch_host_chunks(4096, 0, b)
    assert list(b[: k + 1]) == [0, 148, 444, 1036, 1628, 2220, 2812, 3404, 3652, 3948, 4096]
    assert lib.d3b200_batch_host_chunks(0, 0, None) == 0


def test_module_paths_of_the_reference_exist():
    """Every import path of the reference package resolves here too (drop-in `import tad_dftd3_b200 as tad_dftd3`)."""
    import importlib

    for mod, names in {
        "tad_dftd3_b200": ["dftd3", "damping", "data", "defaults", "disp", "model", "ncoord", "reference", "typing", "__version__"],
        "tad_dftd3_b200.damping.atm": ["dispersion_atm"],
        "tad_dftd3_b200.damping.rational": ["rational_damping"],
        "tad_dftd3_b200.data.r4r2": ["R4R2"],
        "tad_dftd3_b200.model.c6": ["atomic_c6", "_atomic_c6_full", "_atomic_c6_chunked", "AtomicC6_V1", "AtomicC6_V2"],
        "tad_dftd3_b200.model.weights": ["gaussian_weight", "weight_references"],
        "tad_dftd3_b200.ncoord": ["cn_d3", "exp_count"],
        "tad_dftd3_b200.disp": ["dftd3", "dispersion", "dispersion2", "dispersion3"],
        "tad_dftd3_b200.reference": ["Reference"],
        "tad_dftd3_b200.exception": ["DFTD3Error"],
        "tad_dftd3_b200.__version__": ["__version__"],
    }.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n}"


def test_check_memory_contract(monkeypatch):
    """Reference test/test_model/test_util.py: MemoryError above the total, ResourceWarning above the free memory."""
    from tad_dftd3_b200.model import c6

    x = torch.zeros(1000000, dtype=torch.double)
    monkeypatch.setattr(c6, "memory_device", lambda device: (0.0, 0.0))
    with pytest.raises(MemoryError):
        c6._check_memory(x, x)
    monkeypatch.setattr(c6, "memory_device", lambda device: (0.0, 1e10))
    with pytest.warns(ResourceWarning):
        c6._check_memory(torch.zeros(10000, dtype=torch.double), x)
    monkeypatch.undo()
    c6._check_memory(torch.zeros(10, dtype=torch.int64), torch.zeros(10, 7))     # plenty of memory: silent


def test_atm_workspace_holds_the_pair_table_up_to_the_cap(lib, monkeypatch):
    """`d3b200_system_atm_workspace_bytes_*` = neighbour scratch + the dense pair table of the ATM sweep
    (4 reals per atom pair) while the table stays below $D3B200_ATM_PAIRTAB_MAX_BYTES (d3b200.h); no CUDA call."""
    n = 1024
    monkeypatch.setenv("D3B200_ATM_PAIRTAB_MAX_BYTES", "0")
    scratch64 = lib.d3b200_system_atm_workspace_bytes_f64(n)
    scratch32 = lib.d3b200_system_atm_workspace_bytes_f32(n)
    assert scratch64 > 0 and scratch64 % 256 == 0 and scratch32 < scratch64
    monkeypatch.delenv("D3B200_ATM_PAIRTAB_MAX_BYTES")
    assert lib.d3b200_system_atm_workspace_bytes_f64(n) == scratch64 + 32 * n * n
    assert lib.d3b200_system_atm_workspace_bytes_f32(n) == scratch32 + 16 * n * n
    # just below / at the cap
    monkeypatch.setenv("D3B200_ATM_PAIRTAB_MAX_BYTES", str(32 * n * n - 1))
    assert lib.d3b200_system_atm_workspace_bytes_f64(n) == scratch64
    monkeypatch.setenv("D3B200_ATM_PAIRTAB_MAX_BYTES", str(32 * n * n))
    assert lib.d3b200_system_atm_workspace_bytes_f64(n) == scratch64 + 32 * n * n
    # default cap 8 GiB: 16 k atoms get the table, 50 k atoms do not
    monkeypatch.delenv("D3B200_ATM_PAIRTAB_MAX_BYTES")
    big = 50048
    assert lib.d3b200_system_atm_workspace_bytes_f64(big) < 32 * big * big
    assert lib.d3b200_system_atm_workspace_bytes_f64(0) == 0
