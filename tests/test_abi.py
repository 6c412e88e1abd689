"""CPU: the C-ABI library loads and exports exactly what include/aide_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "aide_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aide_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from aide_b200 import _lib
    names = declared_functions()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/aide_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "aide_b200/_lib.py SIGNATURES out of sync with the header"


def test_error_reporting_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    from aide_b200 import _lib
    assert _lib.lib.aide_version() >= 100
    rc = _lib.lib.aide_nchw_to_nhwc(0, None, None, None, 4, 0, 1, 3, 8, 8, None)
    assert rc != 0 and b"bad arguments" in _lib.lib.aide_last_error()
    rc = _lib.lib.aide_conv3x3_fwd(2, 1, None, 24, 0, 24, 1, None, None, 1, 32, 0, 32, 1, 8, 8, None, None)
    assert rc != 0 and b"cin % 32" in _lib.lib.aide_last_error()
    rc = _lib.lib.aide_coteach_select(1, 1, 1, 4, 8, 8, 9, 0.5, 1.0, 10.0, 1.0, 1.0, 1, 1, 1, 1, None, None)
    assert rc != 0 and b"n_clean" in _lib.lib.aide_last_error()


def test_sass_is_blackwell_native():
    """The shipped library contains tcgen05 / TMA / TMEM instructions (SASS mnemonics per B200_PROFILING.md)."""
    import shutil
    import subprocess
    from aide_b200 import _lib
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in sass, mnem
