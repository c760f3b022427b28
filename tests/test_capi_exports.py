"""CPU: the C-ABI shared library loads without a GPU and exports exactly the entry points include/*.h declares;
creating a context without a GPU fails loudly (no CPU fallback)."""
import re
from pathlib import Path

import pytest

from ampliconreconstructorom_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    txt = (ROOT / "include" / "segaligner_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/segaligner_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS and the header disagree"


def test_no_cpu_fallback_without_gpu():
    lib = capi.load_library()
    if lib.sa_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.SegAlignerError):
        capi.Context(0)


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = ROOT / "ampliconreconstructorom_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + \
            list(pkg.rglob("*.cuh")) + [pkg / "csrc" / "Makefile"]:
        txt = f.read_text(errors="ignore")
        assert "sa_oracle" not in txt and "libsa_ref" not in txt and "oracle/" not in txt.replace("oracle/_ref/data", ""), f
