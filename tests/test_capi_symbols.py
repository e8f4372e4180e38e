"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol the
C header declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cwa_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"CWA_API\s+[\w\s\*]+?\b(\w+)\s*\(", txt)))


def test_header_declares_north_star_entry_points():
    syms = declared_symbols()
    for s in ("sph_step", "wave_step", "cwa_coupled_step", "cwa_grid_build", "cwa_scan_exclusive"):
        assert s in syms
    assert len(syms) >= 60


def test_library_builds_and_exports_every_declared_symbol():
    from coupledwateranimation_b200 import build as B
    lib_path = B.build()
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in cwa_b200.h but not exported: {missing}"


def test_python_binding_covers_every_declared_symbol():
    from coupledwateranimation_b200 import _capi
    assert sorted(_capi.SIGNATURES) == declared_symbols()
    _capi.load()


def test_no_device_fails_loudly():
    """Without a CUDA device the context cannot be created: there is no CPU fallback."""
    import coupledwateranimation_b200 as cwa
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(cwa.CwaError) as ei:
        cwa.Context(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    """The product path may mention the oracle in comments but must not import, include, link or
    dlopen anything under oracle/."""
    bad = re.compile(r"(import\s+oracle|from\s+oracle|from\s+\.+oracle|liboracle|cwa_oracle|oracle/|oracle\.py|orc_[a-z0-9_]+\s*\()")
    for top in ("coupledwateranimation_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    src = open(os.path.join(dirpath, f)).read()
                    m = bad.search(src)
                    assert m is None, f"{os.path.join(dirpath, f)} references the oracle: {m.group(0)!r}"
