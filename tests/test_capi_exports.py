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


def test_ctypes_mirrors_have_the_header_layout(tmp_path):
    """Every struct that crosses the C ABI: size and field offsets of the Python mirror equal what gcc makes of the header."""
    import ctypes as C
    import subprocess

    mirrors = {
        "sa_problem": capi.problem_dtype, "sa_result": capi.result_dtype, "sa_round": capi.round_dtype,
        "sa_mode": capi.Mode, "sa_packed_paths": capi.PackedPaths, "sa_rounds_out": capi.RoundsOut,
        "sa_rounds_filter": capi.RoundsFilter, "sa_timing": capi.Timing,
    }

    def fields(m):
        if isinstance(m, capi.np.dtype):
            return [(n, m.fields[n][1]) for n in m.names], m.itemsize
        return [(n, getattr(m, n).offset) for n, _ in m._fields_], C.sizeof(m)

    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/segaligner_b200.h"', "int main(void) {"]
    for name, m in mirrors.items():
        fl, _ = fields(m)
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        for fname, _ in fl:
            lines.append(f'  printf(" %zu", offsetof({name}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {l.split()[0]: [int(v) for v in l.split()[1:]] for l in out if l.strip()}
    for name, m in mirrors.items():
        fl, size = fields(m)
        assert got[name] == [size] + [off for _, off in fl], (name, got[name], size, fl)
