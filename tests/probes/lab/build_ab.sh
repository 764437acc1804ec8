#!/bin/bash
# Build library variants for the A/B probes: tests/probes/lab/build_ab.sh name "-DFLAG=1 ..." [name2 "..."] ...
# SCORE_SRC=tests/probes/lab/pivot_score_r2_poly.cu (or ..._r1_variants.cu) builds the variants from a lab copy of the scoring source.
# -> build/ab/librtk_<name>.so (git-ignored, travels to the GPU box)
set -e
cd "$(dirname "$0")/../../.."
SRC=video-retake_b200/csrc
mkdir -p build/ab
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DNDEBUG -I$SRC $flags \
    -shared -o build/ab/librtk_$name.so $SRC/dpselect.cu $SRC/mallm.cu ${SCORE_SRC:-$SRC/pivot_score.cu} $SRC/pivot_misc.cu -cudart static &
done
wait
ls -la build/ab
