#!/bin/bash
# KA-1 (hg19 BspQI detection sweep on shipped data): wall time of the B200 executable vs the reference binary
set -e
cd "$(dirname "$0")/../.."
D=oracle/_ref/data
mkdir -p /tmp/ka1/new /tmp/ka1/ref
gunzip -c $D/hg19_BspQI.cmap.gz > /tmp/ka1/hg19.cmap
N=$(nproc)
for rep in 1 2 3; do
  t0=$(date +%s.%N)
  SA_VERBOSE=1 ampliconreconstructorom_b200/bin/SegAligner $D/GBM39_amplicon1_graph_BSPQI.cmap /tmp/ka1/hg19.cmap -nthreads=$N -min_labs=10 -prefix=/tmp/ka1/new/SA -detection -gen=1 > /dev/null
  t1=$(date +%s.%N); echo "b200 SegAligner process wall $(python3 -c "print(round($t1-$t0,3))") s (rep $rep)"
done
t0=$(date +%s.%N)
oracle/_ref/SegAligner_ref $D/GBM39_amplicon1_graph_BSPQI.cmap /tmp/ka1/hg19.cmap -nthreads=$N -min_labs=10 -prefix=/tmp/ka1/ref/SA -detection -gen=1 > /dev/null
t1=$(date +%s.%N); echo "reference SegAligner process wall $(python3 -c "print(round($t1-$t0,3))") s (-nthreads=$N)"
for f in /tmp/ka1/ref/*; do cmp -s $f /tmp/ka1/new/$(basename $f) || echo "DIFF $(basename $f)"; done
echo "files: $(ls /tmp/ka1/ref | wc -l) compared"
