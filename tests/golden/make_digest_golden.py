"""Regenerates tests/golden/digest_digests.json by running the UNMODIFIED reference generate_cmap.py
(/root/reference, build container only) on the seeded inputs of tests/digest_cases.py."""
import hashlib
import json
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import digest_cases as dc  # noqa: E402

out = {}
for name, flags in dc.CASES.items():
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        dc.write_inputs(d)
        p = subprocess.run([sys.executable, "/root/reference/generate_cmap.py", "-r", "ref.fa", "-o", "out"] + flags, cwd=d,
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        files = sorted(f for f in d.iterdir() if f.name.startswith("out"))
        out[name] = {f.name: hashlib.sha256(f.read_bytes()).hexdigest() for f in files}
        out[name]["stdout"] = hashlib.sha256(p.stdout.encode()).hexdigest()
        print(name, [f.name for f in files], sum(1 for _ in open(files[0])))
(HERE / "digest_digests.json").write_text(json.dumps(out, indent=0, sort_keys=True))
