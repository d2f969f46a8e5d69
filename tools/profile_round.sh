#!/bin/bash
# Round profile on one B200 (run under gpurun): plain bench, per-launch device times of the same
# command under ncu, and one full-metric capture of the dominant kernel family.  Outputs land in
# gpurun_out/; tools/ncu_summary.py turns the .ncu-rep into the CSV summaries kept in profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 20 -c 12 \
    -o gpurun_out/conv_full_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
