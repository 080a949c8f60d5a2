#!/bin/bash
# Builds A/B variants of libocc_b200.so that differ in compile-time tuning constants of csrc/annotate.cu
# (select one at run time with OCCB200_LIB=<path>): tools/build_variants.sh "name:-DOCC_PPI=8" "name2:-DOCC_MINB=3" ...
set -e
cd "$(dirname "$0")/../objectcentricocccompletion_b200/csrc"
make -s -j8
mkdir -p _build/variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC,-O2,-Wall,-fno-fast-math,-fopenmp -I../../include -I. --expt-relaxed-constexpr"
OTHERS="_build/lib.o _build/points_in_boxes.o _build/voxelize.o _build/scatter.o _build/occ_ops.o _build/range_image.o _build/candidates.o _build/ri_windows.o _build/point_pool.o"
for v in "$@"; do
  name="${v%%:*}"; defs="${v#*:}"
  nvcc $FLAGS $defs -Xptxas -v -c annotate.cu -o _build/variants/annotate_$name.o 2> _build/variants/$name.ptxas.log
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o _build/variants/libocc_b200_$name.so _build/variants/annotate_$name.o $OTHERS -lgomp
  echo "$name: $(grep -A2 'k_visibilityE' _build/variants/$name.ptxas.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'k_visibilityE' _build/variants/$name.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
done
