#!/bin/bash
# Round capture: tests, bench lines of every BASELINE configuration, the PRN sweep, and the ncu evidence.
# usage: bash tools/capture.sh <tag>      (outputs go to gpurun_out/<tag>_*)
set -u
T=${1:-r01e}
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${T}_pytest_gpu.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --impl reference --steps 100 --warmup 5 > $O/${T}_bench_ref.json 2>> $O/${T}_bench.err
for w in c1 c3 c4; do
  python bench.py --workload $w --steps 300 --warmup 20 --no-cpu-baseline > $O/${T}_bench_$w.json 2>> $O/${T}_bench.err
done
python tools/prn_sweep.py 16 78 256 1000 10000 30000 > $O/${T}_prn_sweep.txt 2>&1
python tools/decode_bench.py 78 600 2801 10000 > $O/${T}_decode_bench.txt 2>&1
python tools/ablate.py c2 > $O/${T}_ablate_c2.txt 2>&1
python tools/fused_trace.py 78 2>&1 | tail -15 > $O/${T}_fused_trace.txt
python tools/two_streams.py c2 1 2 3 4 > $O/${T}_lanes.txt 2>&1
python tools/two_streams.py c1 1 3 >> $O/${T}_lanes.txt 2>&1
# ncu: only after the same command has exited 0 without it
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"prn_fused|heatmap_kernel|crop_padded|keypoint_decode|sort_nms|candidates_flat|normalise" \
    --launch-skip 70 --launch-count 7 -f -o $O/${T}_prof_c2 python bench.py --lanes 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_ncu2.log 2>&1
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"big_|crop_padded|keypoint_decode|sort_nms|fc1_reduce" \
    --launch-skip 36 --launch-count 6 -f -o $O/${T}_prof_c3 python bench.py --lanes 1 --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_ncu3.log 2>&1
ls -la $O | grep ${T}_ | awk '{print $5, $9}'
