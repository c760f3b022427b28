#!/usr/bin/env python
"""generate_cmap -- drop-in for /root/reference/generate_cmap.py (in-silico digestion of FASTA / breakpoint-graph
segments / BED regions into a CMAP), SURVEY.md section 8 row f-4.

Same command line, same `<prefix>.cmap` and `<prefix>_key.txt` bytes, same stdout lines.  The motif scan over both
strands (generate_cmap.py:119-123, two regex passes per map in the reference) runs on the GPU through the C ABI
(`sa_digest`); everything else is the file/CLI logic of the reference (read_graph :42-63, read_bed :66-82,
fasta_reader :24-35, makeCMAP :92-132, main :135-208), re-written around byte buffers.  No CPU fallback for the scan.
"""
from __future__ import annotations

import argparse
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ampliconreconstructorom_b200 import capi  # noqa: E402

# enzyme -> (recognition sequence, cut offset)   generate_cmap.py:151-157
ENZYMES = {"BSPQI": ("GCTCTTC", 8), "BBVCI": ("CCTCAGC", 2), "BSMI": ("GAATGC", 4), "BSRDI": ("GCAATG", 5),
           "DLE1": ("CTTAAG", 2)}


def fasta_reader(fasta_file, chroms_to_get, get_all=False):
    """{name: uint8 array}; name = first token of the header; sequence lines are stripped and joined (:24-35)."""
    out = {}
    name, chunks, keep = None, [], False

    def flush():
        if name is not None and keep:
            out[name] = np.frombuffer(b"".join(chunks), dtype=np.uint8)

    with open(fasta_file, "rb") as f:
        for line in f:
            if line[:1] == b">":
                flush()
                name = line[1:].decode().rstrip().rsplit()[0]
                keep = get_all or (name in chroms_to_get)
                chunks = []
                if keep:
                    print("Reading " + name)
            elif keep:
                chunks.append(line.strip())
        flush()
    return out


def read_graph(graph_file):
    segs, chroms = [], set()
    with open(graph_file) as f:
        next(f)
        for line in f:
            if line.startswith("sequence"):
                fields = line.rstrip().split()
                f1, f2 = fields[1].rsplit(":"), fields[2].rsplit(":")
                chroms.add(f1[0])
                p1, p2 = int(f1[1][:-1]), int(f2[1][:-1])
                segs.append((fields[1] + "|" + fields[2], f1[0], min(p1, p2) - 1, max(p1, p2) - 1))
    return segs, chroms


def read_bed(bed_file):
    bed = defaultdict(list)
    with open(bed_file) as f:
        for line in f:
            if not line.startswith("#"):
                fields = line.rstrip().rsplit()
                bed[fields[0]].append((int(fields[1]), int(fields[2])))
    segs = []
    for chrom, regions in bed.items():
        for p in regions:
            segs.append((str(p[0]) + "-|" + str(p[1]) + "+", chrom, p[0] - 1, p[1] - 1))
    return segs, set(bed.keys())


def make_cmap(ctx, outprefix, output_arg, segs, seg_seq, enzyme, min_label, min_size):
    print("Generating CMAP for enzyme " + enzyme)
    motif, offset = ENZYMES[enzyme]
    offset += 1  # "cmap 1 based" (:102)
    lines = ["# CMAP File Version:\t0.1\n# Label Channels:\t1\n# Nickase Recognition Site 1:\t" + enzyme +
             "\n# Enzyme1:\t" + enzyme + "\n# Number of Consensus Nanomaps:\t" + str(len(seg_seq)) +
             "\n#h CMapId\tContigLength\tNumSites\tSiteID\tLabelChannel\tPosition\tStdDev\tCoverage\tOccurrence\n"
             "#f int\tfloat\tint\tint\tint\tfloat\tfloat\tint\tint\n"]
    keys = ["# CMAP = " + output_arg + "_key.txt\n# filter: Minimum Labels = " + str(min_label) +
            "\n# filter: Minimum Size (Kb) = " + str(min_size / 1000.0) + "\n"]
    for num, seg in enumerate(segs):
        name = seg[0]
        cmap_id = str(num + 1)
        seq = seg_seq[name]
        map_len = float(len(seq))
        if map_len < min_size:
            continue
        map_len_s = str(map_len)
        print("Generating map for " + name)
        pos = ctx.digest(seq, motif, offset)  # both strands, sorted, de-duplicated  (:119-123)
        if len(pos) < min_label:
            continue
        total = str(len(pos))
        body = [cmap_id + "\t" + map_len_s + "\t" + total + "\t" + str(i + 1) + "\t1\t" + str(float(v)) + "\t1.0\t1\t1\n"
                for i, v in enumerate(pos.tolist())]
        lines.extend(body)
        lines.append("\t".join([cmap_id, map_len_s, total, str(len(pos) + 1), "0", map_len_s, "1.0\t1\t1\n"]))
        keys.append("\t".join([cmap_id, name, map_len_s]) + "\n")
    with open(outprefix + ".cmap", "w") as f:
        f.write("".join(lines))
    with open(outprefix + "_key.txt", "w") as f:
        f.write("".join(keys))
    print("Finished")


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-r", "--ref", help="reference genome fasta", required=True)
    ap.add_argument("-o", "--output", help="output prefix")
    ap.add_argument("-e", "--enzyme", help="restriction enzyme: BspQI, BbvCI, BsmI, BsrDI, DLE1", required=True)
    ap.add_argument("-l", "--labels", type=int, help="minimum number of labels in reported CMAP, default 0", default=0)
    ap.add_argument("-s", "--size", type=float, help="minimum map size for reported CMAP (basepairs), default 0", default=0)
    group = ap.add_mutually_exclusive_group(required=True)
    group.add_argument("-g", "--graph", help="breakpoint graph .txt file, if not supplied, --makeRef or --bed must be supplied")
    group.add_argument("--makeRef", help="Make CMAP from entire reference fasta", action="store_true")
    group.add_argument("-b", "--bed", help="breakpoint graph bed file, it not supplied, --makeRef or --graph must be supplied")
    args = ap.parse_args(argv)
    if args.labels < 0:
        sys.exit("Innaproriate min number of labels " + str(args.labels) + ", exiting")
    if args.size < 0:
        sys.exit("Innaproriate min map size " + str(args.size) + ", exiting")
    args.enzyme = args.enzyme.upper()
    if args.enzyme not in ENZYMES:
        print("Valid enzymes are " + str(ENZYMES.keys()))
        sys.exit("Unrecognized enzyme " + args.enzyme + ", exiting")
    if not args.output:
        if args.graph:
            args.output = ".".join(args.graph.rsplit(".")[:-1])
        elif args.bed:
            args.output = ".".join(args.bed.rsplit(".")[:-1])
        else:
            args.output = ".".join(args.ref.rsplit("/")[-1].rsplit(".")[:-1])
    outprefix = args.output
    if args.enzyme not in args.output:
        outprefix = outprefix + "_" + args.enzyme

    if args.graph:
        segs, chroms = read_graph(args.graph)
        seqs = fasta_reader(args.ref, chroms)
    elif args.bed:
        segs, chroms = read_bed(args.bed)
        seqs = fasta_reader(args.ref, chroms)
    else:
        seqs = fasta_reader(args.ref, set(), True)
        segs = []
        for name in sorted(seqs, key=lambda x: x.lstrip("chr")):
            if len(seqs[name]) > args.size:
                segs.append((name, name, 0, len(seqs[name])))
    seg_seq = {}
    for s in segs:  # segsToSeq (:85-90): python slice semantics (negative / oversized bounds clamp like str slicing)
        seg_seq[s[0]] = seqs[s[1]][slice(s[2], s[3])]
    ctx = capi.Context(0)  # raises without an sm_100 GPU: the scan has no CPU fallback
    make_cmap(ctx, outprefix, args.output, segs, seg_seq, args.enzyme, args.labels, args.size)
    ctx.close()


if __name__ == "__main__":
    main()
