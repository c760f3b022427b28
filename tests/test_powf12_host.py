"""CPU: the powf(x,1.2f) replay (csrc/powf12.cuh compiled as host C++) against the libm powf the reference links,
EXHAUSTIVELY over every non-negative float bit pattern incl. zero, subnormals, inf and the NaN boundary
(2.1e9 values; ~45 CPU-seconds).  The device executes the same IEEE-754 operation sequence (DFMA = fused),
and tests/test_gpu_parity.py re-checks both device paths against libm on the GPU."""
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "ampliconreconstructorom_b200" / "csrc" / "tools" / "powf_check_host.cpp"


def test_powf12_replay_is_bit_exact_for_all_floats(tmp_path):
    exe = tmp_path / "powf_check_host"
    subprocess.check_call(["g++", "-O2", "-mfma", "-std=c++17", "-pthread", "-o", str(exe), str(SRC), "-lm"])
    n = str(os.cpu_count() or 1)
    out = subprocess.run([str(exe), "0", "0x7f800010", n], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches=0" in out.stdout
