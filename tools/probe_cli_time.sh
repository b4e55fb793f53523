#!/bin/bash
# Developer probe: where does the wall time of `deltapq -task approx_tree` go in a fresh process?
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
import datagen as dg
dg.make_dataset("$D", 1000000, 100, M=8, K=256, d=128, seed=0, n_learn=20000)
PY
B=$ROOT/deltapq_b200/bin
C="-dataset $D -m 8 -k 256 -N 1000000 -ext fvecs"
T0=$(date +%s%N); $B/pqtree -task encode $C > /dev/null; echo "encode $(( ($(date +%s%N) - T0) / 1000000 )) ms"
for i in 1 2; do
  rm -f $D/M8K256*_Approx_*
  T0=$(date +%s%N); $B/deltapq -task approx_tree -h 1 -diff 8 $C | tail -2; echo "approx_tree run $i: $(( ($(date +%s%N) - T0) / 1000000 )) ms"
done
rm -f $D/M8K256*_Approx_*
T0=$(date +%s%N); DPQ_HOST_LAYOUT=1 $B/deltapq -task approx_tree -h 1 -diff 8 $C | tail -1; echo "approx_tree host layout: $(( ($(date +%s%N) - T0) / 1000000 )) ms"
rm -f $D/M8K256*_Approx_*
T0=$(date +%s%N); CUDA_MODULE_LOADING=EAGER $B/deltapq -task approx_tree -h 1 -diff 8 $C | tail -1; echo "approx_tree eager: $(( ($(date +%s%N) - T0) / 1000000 )) ms"
ls -la $D | head -12
rm -rf $D
