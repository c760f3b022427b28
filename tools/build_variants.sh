#!/bin/bash
# Developer tool: build knock-out / knob variants of libsegaligner_b200.so into tools/variants/<name>/ (git-ignored)
# for A/B timing with `SA_LIB=tools/variants/<name>/libsegaligner_b200.so python tools/quick_bench.py ...`.
# usage: tools/build_variants.sh name1:"-DFLAG1 -DFLAG2" name2:"..."
cd "$(dirname "$0")/../ampliconreconstructorom_b200/csrc"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  out=../../tools/variants/$name; mkdir -p $out
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2,-ffp-contract=off -Xptxas -v $flags \
      -I../../include -shared -o $out/libsegaligner_b200.so sa_kernels.cu sa_capi.cu sa_digest.cu host/cmap.cpp host/host_capi.cpp 2> $out/ptxas.log \
      && echo "built $name" || (echo "FAILED $name"; tail -5 $out/ptxas.log) ) &
done
wait
