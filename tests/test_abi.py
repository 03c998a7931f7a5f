"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/b200cv.h declares; no compute call is made (there is no GPU here), and the
product refuses to run without one instead of falling back to the CPU."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    hdr = (ROOT / "include" / "b200cv.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200cv_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    import colvars_b200 as cb
    assert declared_symbols() == sorted(cb.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import colvars_b200 as cb
    if not cb.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(str(cb.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    import torch
    import colvars_b200 as cb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.B200cvError, match="no CPU fallback"):
        cb.CoordNum("coordNum", [0, 1], [2, 3])


def test_product_never_touches_the_oracle():
    """Nothing under colvars_b200/ or include/ may import, link or name oracle/."""
    for path in list((ROOT / "colvars_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if path.is_file() and path.suffix in (".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ""):
            if path.name.startswith("Makefile") or path.suffix:
                txt = path.read_text(errors="ignore")
                assert "oracle" not in txt.replace("oracle/_ref", "").lower() or \
                    path.name == "INTEGRATION.md", path
