"""The two steps either side of the hg19 detection sweep: drop-in replacements for

  get_unaligned_segs            /root/reference/ARAlignDetect.py:31-78    unaligned label runs per contig
  write_unaligned_cmaps         /root/reference/ARAlignDetect.py:83-117   region CMAP written for `SegAligner -detection`
  write_aligned_labels          /root/reference/ARAlignDetect.py:137-155
  detections_to_seg_alignments  /root/reference/ARAlignDetect.py:198-280  detection alignments -> graph segments + *_rg_aln.txt

Same names, same arguments, same return values, byte-identical files.  The reference versions read the module globals
`contig_cmaps` / `contig_lens` that ARAlignDetect's main sets (ARAlignDetect.py:359-361); here they are module globals too
(set them once: `align_detect.contig_cmaps = parse_cmap(...)`), or pass them as keyword arguments.  The label bookkeeping
is done on numpy bitmaps (one pass per contig) instead of Python set arithmetic over every label; the parsers mirror
bionanoUtil.py (parse_cmap :12-35, get_cmap_lens :38-51, parse_keyfile :89-100, parse_seg_alignment_file :335-357).
Host-side Python like the reference's own code for this row; nothing here touches the GPU.
"""
from __future__ import annotations

import os
from collections import defaultdict

import numpy as np

# ARAlignDetect.py:25-28
unaligned_label_cutoff = 20
search_upper_cutoff_labels = 500
unaligned_size_lower_cutoff = 200000
unaligned_size_upper_cutoff = 5000000

contig_cmaps: dict = {}
contig_lens: dict = {}


# ---- parsers (bionanoUtil.py) ----------------------------------------------------------------------------------------
def parse_cmap(cmapf, keep_length=False):
    """bionanoUtil.py:12-35: {CMapId(str): {SiteID(int): Position(float)}}, label channel 1 only unless keep_length."""
    cmaps = {}
    head = None
    with open(cmapf) as infile:
        for line in infile:
            if line.startswith("#h"):
                head = line.rstrip().rsplit()[1:]
            elif not line.startswith("#"):
                fD = dict(zip(head, line.rstrip().rsplit()))
                m = cmaps.setdefault(fD["CMapId"], {})
                if fD["LabelChannel"] == "1" or (fD["LabelChannel"] == "0" and keep_length):
                    m[int(fD["SiteID"])] = float(fD["Position"])
    return cmaps


def get_cmap_lens(cmapf):
    """bionanoUtil.py:38-51."""
    lens = {}
    head = None
    with open(cmapf) as infile:
        for line in infile:
            if line.startswith("#h"):
                head = line.rstrip().rsplit()[1:]
            elif not line.startswith("#"):
                fD = dict(zip(head, line.rstrip().rsplit()))
                if fD["CMapId"] not in lens:
                    lens[fD["CMapId"]] = float(fD["ContigLength"])
    return lens


def parse_keyfile(keyF_name):
    """bionanoUtil.py:89-100: {CompntId: (CompntName, CompntLength)}."""
    out = {}
    with open(keyF_name) as infile:
        for line in infile:
            if not line.startswith("#") and not line.startswith("CompntId"):
                f = line.rstrip().split()
                out[f[0]] = (f[1], float(f[2]))
    return out


def parse_aln_endpoints(alignfile):
    """The part of parse_seg_alignment_file (bionanoUtil.py:335-357) the region extraction uses: contig id and the
    (first, last) contig labels of the alignment rows (1-based label numbers, as SegAligner writes them)."""
    with open(alignfile) as infile:
        next(infile)
        next(infile)
        aln_head = next(infile).rstrip()[1:].rsplit()
        ci, li = aln_head.index("contig_id"), aln_head.index("contig_label")
        first = last = None
        for line in infile:
            fields = line.rstrip().rsplit()
            if first is None:
                first = fields
            last = fields
    return first[ci], (int(first[li]), int(last[li]))


# ---- ARAlignDetect.py:31-78 ------------------------------------------------------------------------------------------
def get_unaligned_segs(aln_path, aln_flist, contig_cmaps=None):
    """Per contig (in order of first appearance in aln_flist): the runs of consecutive labels no alignment covers, kept
    when a run spans more than 20 labels and more than 200 kb; runs over 5 Mb are cut to their first 500 labels; contigs
    that are more than 90 % unaligned are skipped.  Label bookkeeping as in the reference: the free set is 0..n-1 while
    alignment end points are the 1-based label numbers of the files (ARAlignDetect.py:44-48)."""
    cm = globals()["contig_cmaps"] if contig_cmaps is None else contig_cmaps
    ends_by_contig = defaultdict(list)
    for f in aln_flist:
        c_id, ends = parse_aln_endpoints(aln_path + f)
        ends_by_contig[c_id].append(ends)
    regions = defaultdict(list)
    for c_id, ends in ends_by_contig.items():
        labels = cm[c_id]
        n = len(labels)
        free = np.ones(n, dtype=bool)
        for a, b in ends:  # set(range(a, b + 1)) removed from set(range(n))
            lo, hi = max(a, 0), min(b + 1, n)
            if hi > lo:
                free[lo:hi] = False
        if float(int(free.sum())) / n > 0.9:
            continue
        idx = np.flatnonzero(free)
        if idx.size == 0:
            continue
        cuts = np.flatnonzero(np.diff(idx) > 1) + 1
        for run in np.split(idx, cuts):
            first, last = int(run[0]), int(run[-1])
            size = labels[last + 1] - labels[first + 1]
            if last - first > unaligned_label_cutoff and size > unaligned_size_lower_cutoff:
                l = [int(x) for x in run]
                regions[c_id].append(l[:search_upper_cutoff_labels] if size > unaligned_size_upper_cutoff else l)
    return regions


# ---- ARAlignDetect.py:83-117 -----------------------------------------------------------------------------------------
def write_unaligned_cmaps(contig_unaligned_regions, output_prefix, enzyme, contig_cmaps=None, contig_lens=None):
    """Writes <output_prefix>contig_unaligned_regions.cmap (one map per region, ids continuing after the largest contig
    id, positions relative to the region's first label, floats printed by str()) and returns
    (new id -> {region label: contig label}, file name, new id -> contig id)."""
    cm = globals()["contig_cmaps"] if contig_cmaps is None else contig_cmaps
    cl = globals()["contig_lens"] if contig_lens is None else contig_lens
    max_contig_id = max([int(x) for x in cm])
    label_trans, id_to_contig = {}, {}
    nmaps = sum([len(x) for x in contig_unaligned_regions.values()])
    cmap_header = ("# CMAP File Version: 0.1\n# Label Channels:  1\n# Nickase Recognition Site 1:    " + enzyme +
                   "\n# Enzyme1:  " + enzyme + "\n# Number of Consensus Nanomaps: " + str(nmaps) +
                   "\n#h CMapId ContigLength    NumSites    SiteID  LabelChannel    Position    StdDev  Coverage    "
                   "Occurrence\n#f int  float   int int int float   float   int int\n")
    fname = output_prefix + "contig_unaligned_regions.cmap"
    out = [cmap_header]
    total = 0
    for c_id, all_regions in contig_unaligned_regions.items():
        labels = cm[c_id]
        for l in all_regions:
            total += 1
            map_end_pos = labels[l[-1] + 2] if (l[-1] + 2) in labels else cl[c_id]
            pos_l = [labels[x + 1] for x in l]
            curr_id = str(max_contig_id + total)
            label_trans[curr_id] = {str(ind + 1): str(x + 1) for ind, x in enumerate(l)}
            id_to_contig[curr_id] = c_id
            p0 = pos_l[0]
            map_len_str = str(map_end_pos - p0)
            n_str = str(len(l))
            for ind, p in enumerate(pos_l):
                out.append("\t".join([curr_id, map_len_str, n_str, str(ind + 1), "1", str(p - p0), "1.0", "1", "1"]) + "\n")
            # the end row repeats the last SiteID (the reference reuses `ind`, ARAlignDetect.py:114)
            out.append("\t".join([curr_id, map_len_str, n_str, str(len(pos_l)), "0", map_len_str, "1.0", "1", "1"]) + "\n")
    with open(fname, "w") as outfile:
        outfile.write("".join(out))
    return label_trans, fname, id_to_contig


# ---- ARAlignDetect.py:137-155 ----------------------------------------------------------------------------------------
def write_aligned_labels(a_dir, aln_flist, w_dir):
    covered = defaultdict(set)
    for f in aln_flist:
        c_id, ends = parse_aln_endpoints(a_dir + f)
        covered[c_id] |= set(range(min(ends), max(ends) + 1))
    outfile_name = w_dir + "contig_aligned_labels_nontip.txt"
    with open(outfile_name, "w") as outfile:
        outfile.write("#contig_id [space delimited list of labels]\n")
        for c_id in covered:
            outfile.write(" ".join([c_id] + [str(x - 1) for x in sorted(covered[c_id])]) + "\n")
    return outfile_name


# ---- ARAlignDetect.py:198-280 ----------------------------------------------------------------------------------------
def detections_to_seg_alignments(w_dir, aln_files, ref_file, unaligned_cid_d, unaligned_label_trans, id_start, prefix):
    """Turns the `SegAligner -detection` alignments of the unaligned regions against the reference genome into new graph
    segments: returns bpg_list = [(chrom, start-10, end+10)] in order of discovery and writes, per detection alignment,
    <prefix>_<contig>_<new segment id>_rg[_r]_aln.txt with the region labels translated back to contig labels."""
    ref_genome_cmaps = parse_cmap(ref_file)
    ref_genome_key_dict = parse_keyfile(os.path.splitext(ref_file)[0] + "_key.txt")
    bpg_list = []
    u_id_lookup = {}
    aln_num = 0
    print("aln_files len", len(aln_files))
    for f in aln_files:
        with open(w_dir + f) as infile:
            head_list = [next(infile) for _ in range(3)]
            aln_field_list = [line.rstrip().rsplit("\t") for line in infile]
        meta = head_list[1].rsplit("\t")
        contig_id = meta[0][1:-1]  # in detection runs the "segment" of the file is the unaligned region (a contig piece)
        contig_dir = meta[0][-1]
        if contig_id not in unaligned_cid_d:
            continue
        trans_contig_id = unaligned_cid_d[contig_id]
        ref_genome_id = aln_field_list[0][0]
        chromID = ref_genome_key_dict[ref_genome_id][0]
        p1 = int(ref_genome_cmaps[ref_genome_id][int(aln_field_list[0][2])])
        p2 = int(ref_genome_cmaps[ref_genome_id][int(aln_field_list[-1][2])])
        if p1 > p2:
            p1, p2 = p2, p1
        key = (chromID, str(p1 - 10), str(p2 + 10))  # 10 bp padding on the aligned segment
        if key not in u_id_lookup:
            bpg_list.append(key)
            aln_num += 1
            u_id_lookup[key] = str(id_start + aln_num)
        unique_id = u_id_lookup[key]
        if contig_dir == "-":
            aln_field_list = aln_field_list[::-1]
        meta_string = "#" + "\t".join([unique_id + contig_dir, meta[1], "False"]) + "\n"
        is_rev = "_r" if contig_dir == "-" else ""
        outname = prefix + "_" + trans_contig_id + "_" + unique_id + "_rg" + is_rev + "_aln.txt"
        rows = [head_list[0], meta_string, head_list[2]]
        curr_score = 0.0
        lchange = 0
        first = min(int(aln_field_list[0][2]), int(aln_field_list[-1][2]))
        for l in aln_field_list:
            trans_lab = unaligned_label_trans[contig_id][l[3]]
            rows.append("\t".join([trans_contig_id, unique_id, trans_lab, str(int(l[2]) - first + 1)] + l[4:7] +
                                  [str(curr_score), str(lchange)]) + "\n")
            curr_score += float(l[-1])
            lchange = l[-1]
        with open(w_dir + outname, "w") as outfile:
            outfile.write("".join(rows))
    return bpg_list
