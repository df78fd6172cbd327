#!/bin/bash
# Round capture: tests, bench lines of every BASELINE configuration, the PRN sweep, and the ncu evidence.
# usage: bash tools/capture.sh <tag> [stages]     stages: any of t(ests) b(ench) c(onfigs) s(weeps) n(cu) x(sanitizer); default tbcsn
# outputs go to gpurun_out/<tag>_*
set -u
T=${1:-r02a}
S=${2:-tbcsn}
O=gpurun_out
mkdir -p $O
has() { [[ "$S" == *"$1"* ]]; }
if has t; then
  ( time python -m pytest tests -m gpu -q --durations=15 ) > $O/${T}_pytest_gpu.log 2>&1
  tail -25 $O/${T}_pytest_gpu.log
fi
if has b; then
  python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
  python bench.py --steps 20 --warmup 5 > $O/${T}_bench_driver_style.json 2>> $O/${T}_bench.err
  python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_ref.json 2>> $O/${T}_bench.err
  python tools/ablate.py c2 > $O/${T}_ablate_c2.txt 2>&1
  python tools/fused_trace.py 78 2>&1 | tail -15 > $O/${T}_fused_trace.txt
fi
if has c; then
  for w in c1 c3 c4; do
    python bench.py --workload $w --steps 300 --warmup 20 --no-cpu-baseline > $O/${T}_bench_$w.json 2>> $O/${T}_bench.err
  done
  python bench.py --prn-mode fp32 --steps 200 --warmup 20 --no-cpu-baseline > $O/${T}_bench_fp32.json 2>> $O/${T}_bench.err
fi
if has s; then
  python tools/prn_sweep.py 16 78 256 1000 10000 30000 100000 > $O/${T}_prn_sweep.txt 2>&1
  python tools/decode_bench.py 78 600 2801 10000 > $O/${T}_decode_bench.txt 2>&1
  python tools/two_streams.py c2 1 2 3 4 > $O/${T}_lanes.txt 2>&1
  { echo "== crop_and_resize inside the PRN kernel (MPN_FUSE_CROP=1)"; MPN_FUSE_CROP=1 python tools/two_streams.py c2 1 3 2>&1 | tail -2;
    MPN_FUSE_CROP=1 python tools/fused_trace.py c2 2>&1 | tail -17; echo "== default"; python tools/fused_trace.py c2 2>&1 | tail -12;
    echo "== sort / NMS phases"; python tools/nms_trace.py c2 2>&1 | tail -7; python tools/nms_trace.py c3 2>&1 | tail -7; } > $O/${T}_traces.txt 2>&1
fi
if has n; then
  # ncu: only after the same command has exited 0 without it
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --min-seconds 0 > $O/${T}_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu-baseline --min-seconds 0 > $O/${T}_ncu1.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"prn_fused|heatmap_norm|logit_minmax|crop_padded|keypoint_decode|sort_nms|candidates_flat" \
      --launch-skip 70 --launch-count 7 -f -o $O/${T}_prof_c2 python bench.py --lanes 1 --steps 5 --warmup 3 --no-cpu-baseline --min-seconds 0 > $O/${T}_ncu2.log 2>&1
  python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --min-seconds 0 > $O/${T}_plain_c3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"big_|crop_padded|keypoint_decode|sort_nms|fc1_reduce|heatmap_norm|logit_minmax" \
      --launch-skip 40 --launch-count 8 -f -o $O/${T}_prof_c3 python bench.py --lanes 1 --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --min-seconds 0 > $O/${T}_ncu3.log 2>&1
fi
if has x; then
  bash tools/sanitize.sh $T
fi
ls -la $O | grep ${T}_ | awk '{print $5, $9}'
