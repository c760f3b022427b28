"""CPU: bench.py reads the executable's SA_VERBOSE report (phase times, copy counters, GPUs used) without a GPU."""
import importlib.util
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

SAMPLE = """SegAligner(b200): parse + prepare maps            0.203 s
SegAligner(b200): open GPU context                0.712 s
  rank 2 scored 8 units, 4.69e+10 of 3.01e+11 cells, done 6.745 s after start
SegAligner(b200): phase 1 scoring (GPU)           5.988 s
  sa_align_rounds chunk: 6771 pairs, kernel 254.09 ms, 359766 path entries, 3552 workers x 7.3 MB scratch
SegAligner(b200): phase 3 alignment               1.069 s
SegAligner(b200): phase 4 tip alignment           0.317 s
SegAligner(b200): copied H2D 420433736 bytes, D2H 57767924 bytes
SegAligner(b200): wall 8.422 s on 4 GPU(s)
SegAligner(b200): close                           0.754 s
"""


def test_parse_verbose():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    phases, h2d, d2h, used = bench.parse_verbose(SAMPLE)
    assert phases == {"parse + prepare maps": 0.203, "open GPU context": 0.712, "phase 1 scoring (GPU)": 5.988,
                      "phase 3 alignment": 1.069, "phase 4 tip alignment": 0.317, "close": 0.754}
    assert (h2d, d2h, used) == (420433736, 57767924, 4)
    assert bench.parse_verbose("")[3] is None
