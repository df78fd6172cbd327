#!/bin/bash
# compute-sanitizer over the tiny-size GPU tests (every kernel of the path, both PRN modes, host path, lanes):
#   bash tools/sanitize.sh <tag>      -> gpurun_out/<tag>_sanitizer_{memcheck,racecheck,synccheck,initcheck}.txt
# The tools slow kernels down 10-100x, so only tests at the smallest sizes are selected.
set -u
T=${1:-r02}
O=gpurun_out
mkdir -p $O
SEL='tiny or edge or known_answers or no_person or near_ties or graph_replay or pinned_inputs'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_internal_parity.py -m gpu -q -x -k "$SEL" \
      > $O/${T}_sanitizer_${tool}.log 2>&1
  rc=$?
  { echo "tool=$tool exit=$rc"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $O/${T}_sanitizer_${tool}.log | tail -8; } \
      > $O/${T}_sanitizer_${tool}.txt
  cat $O/${T}_sanitizer_${tool}.txt
done
