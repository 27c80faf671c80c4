"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/ctr_b200.h declares, and rejects bad arguments without touching a GPU."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ctr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from tf_repos_b200 import _lib
    names = _declared_symbols()
    assert "ctr_fm_embed_fwd" in names and len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ctr_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in tf_repos_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names)
    assert _lib.abi_version() == 1


def test_argument_validation_needs_no_gpu():
    from tf_repos_b200 import _lib
    L = _lib.raw()
    # bad id width / missing buffers are rejected before any CUDA call
    assert L.ctr_fm_embed_fwd(None, 16, None, None, None, 10, 1, 1, 4, 0, None, None, None, None, None, None) == -1
    assert "id_bits" in _lib.last_error()
    assert L.ctr_fm_embed_fwd(None, 32, None, None, None, 10, 1, 1, 4, 0, None, None, None, None, None, None) == -1
    assert L.ctr_unique_segment(None, 10, 0, None, None, None, None, None, None, None, 0, None) == -1
    assert L.ctr_opt_sparse_rows(9, None, None, None, None, None, None, 5, 4, None, None, None) == -1
    assert L.ctr_unique_segment_workspace_bytes(319488, 200_000_000) > 3 * 319488 * 4
    assert L.ctr_launch_count() == 0


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib
    from tf_repos_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib._load()
    except _lib.CtrError as e:
        assert "no CPU/eager fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
