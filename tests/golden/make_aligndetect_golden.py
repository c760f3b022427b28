"""Regenerates tests/golden/align_detect.json from the UNMODIFIED reference (build container only: needs
/root/reference and oracle/_ref/SegAligner_ref).

The fixture pins SURVEY.md section 8 row f-3 -- ARAlignDetect.py's get_unaligned_segs / write_unaligned_cmaps /
write_aligned_labels / detections_to_seg_alignments -- by IMPORTING the reference module (matplotlib, which this image
lacks and these functions never touch, is stubbed) and calling its own functions on a seeded scenario:
  1. a small in-silico genome of 24 chromosomes (CMAP + _key.txt), graph segments cut from it, contigs that are noisy
     windows around those segments and extend far beyond them (large unaligned flanks, one > 5 Mb);
  2. the reference SegAligner aligns segments to contigs (standard mode)            -> *_aln.txt       (inputs)
  3. reference get_unaligned_segs + write_unaligned_cmaps + write_aligned_labels    -> expected outputs
  4. the reference SegAligner searches the regions against the genome (-detection)  -> *_ref_search_*_aln.txt (inputs)
  5. reference detections_to_seg_alignments                                          -> expected outputs
The alignment files (which only the reference binary can produce) and every expected output are stored in the JSON
(compressed); the CMAP inputs are regenerated from the seed (tests/aligndetect_case.py) and pinned by their sha256.  The CPU
test needs neither the reference nor a GPU.
"""
import io
import json
import os
import subprocess
import sys
import tempfile
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE.parent))
from ampliconreconstructorom_b200 import synth  # noqa: E402
import cli_cases as cc  # noqa: E402
from aligndetect_case import scenario, pack  # noqa: E402

REF = Path("/root/reference")


def import_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, str(REF))
    import ARAlignDetect  # noqa: E402  (the reference module, unmodified)
    return ARAlignDetect


def read_dir(d: Path, pred):
    return {f.name: f.read_text() for f in sorted(d.iterdir()) if pred(f.name)}


def main():
    AD = import_reference()
    out = {}
    with tempfile.TemporaryDirectory() as t:
        tmp = Path(t)
        scenario(tmp)
        a_dir = tmp / "alignments"
        a_dir.mkdir()
        subprocess.run([str(cc.REF_BIN), str(tmp / "segs.cmap"), str(tmp / "contigs.cmap"), "-nthreads=4", "-min_labs=10",
                        f"-prefix={a_dir}/SA_segs", "-gen=1"], stdout=subprocess.DEVNULL, check=True)
        aln_flist = sorted(x for x in os.listdir(a_dir) if "_aln.txt" in x and "flipped" not in x and "_ref_" not in x and "rg" not in x)
        assert aln_flist, "scenario produced no alignments"
        # the CMAP inputs are regenerated from the seed at test time (tests/aligndetect_case.py); their hashes pin them
        import hashlib
        out["input_sha256"] = {n: hashlib.sha256((tmp / n).read_bytes()).hexdigest() for n in ("contigs.cmap", "genome.cmap", "genome_key.txt")}
        out["aln_flist"] = aln_flist
        out["aln_files"] = {f: (a_dir / f).read_text() for f in aln_flist}
        # --- reference functions (module globals as ARAlignDetect's main sets them, :359-361)
        AD.contig_cmaps = AD.parse_cmap(str(tmp / "contigs.cmap"))
        AD.contig_lens = AD.get_cmap_lens(str(tmp / "contigs.cmap"))
        regions = AD.get_unaligned_segs(str(a_dir) + "/", aln_flist)
        assert regions, "scenario produced no unaligned regions"
        out["regions"] = {c: [list(map(int, l)) for l in v] for c, v in regions.items()}
        out["regions_order"] = list(regions.keys())
        trans, fname, cid_d = AD.write_unaligned_cmaps(regions, str(tmp) + "/", "BspQI")
        out["label_trans"], out["cid_d"] = trans, cid_d
        out["regions_cmap"] = Path(fname).read_text()
        lab_file = AD.write_aligned_labels(str(a_dir) + "/", aln_flist, str(tmp) + "/")
        out["aligned_labels"] = Path(lab_file).read_text()
        # --- detection sweep of the regions against the genome, then the reference post-processing
        # run_SegAligner(ref_genome_file, unaligned_region_filename, ...) puts the REGIONS first (ARAlignDetect.py:283-287,411)
        subprocess.run([str(cc.REF_BIN), fname, str(tmp / "genome.cmap"), "-nthreads=4", "-min_labs=10",
                        f"-prefix={a_dir}/SA_segs_ref_search", "-detection", "-gen=1"], stdout=subprocess.DEVNULL, check=True)
        det_flist = sorted(x for x in os.listdir(a_dir) if "_aln.txt" in x and "SA_segs_ref_search" in x)
        assert det_flist, "scenario produced no detection alignments"
        out["det_flist"] = det_flist
        out["det_files"] = {f: (a_dir / f).read_text() for f in det_flist}
        before = set(os.listdir(a_dir))
        buf = io.StringIO()
        with redirect_stdout(buf):
            bpg = AD.detections_to_seg_alignments(str(a_dir) + "/", det_flist, str(tmp / "genome.cmap"), cid_d, trans, 14, "SA_segs")
        out["bpg_list"] = [list(x) for x in bpg]
        out["det_stdout"] = buf.getvalue()
        out["rg_files"] = {f: (a_dir / f).read_text() for f in sorted(set(os.listdir(a_dir)) - before)}
        assert out["rg_files"], "no _rg_ alignment files written"
    (HERE / "align_detect.json").write_text(json.dumps({"format": "zlib+base64 of the JSON document (tests/aligndetect_case.py: unpack)",
                                                        "blob": pack(out)}))
    print(f"{len(out['aln_flist'])} alignments, regions per contig {[(c, [len(l) for l in v]) for c, v in out['regions'].items()]}, "
          f"{len(out['det_flist'])} detection alignments, {len(out['rg_files'])} rg files, bpg {out['bpg_list']}; "
          f"{(HERE / 'align_detect.json').stat().st_size} bytes")


if __name__ == "__main__":
    main()
