#!/bin/bash
# Round profile on one B200 (run under gpurun after the same commands exited 0 without ncu):
#   1. plain bench (never a number from a profiled run),
#   2. per-launch device times of the profiled command (launch list),
#   3. DRAM bytes per kernel of ONE step with the caches left alone (tools/step_traffic.py),
#   4. one --set full capture of the conv kernel families (C1) and of the F = 512 kernels (C4).
# Outputs land in gpurun_out/; tools/ncu_summary.py and tools/step_traffic.py turn them into the
# summaries kept in profiles/.
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
for CFG in c1 c4; do
  python tools/step_traffic.py run $CFG > /dev/null 2>&1 || { echo "step_traffic $CFG failed"; continue; }
  ncu --profile-from-start off --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --csv --log-file gpurun_out/step_traffic_${CFG}.csv python tools/step_traffic.py run $CFG \
      > gpurun_out/ncu_traffic_${CFG}.log 2>&1
  echo "step traffic $CFG rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad3_tc_kernel" \
    --profile-from-start off -c 24 -o gpurun_out/conv_full_${TAG} python tools/step_traffic.py run c1 \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture c1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad3_tc_kernel" \
    --profile-from-start off -c 40 -o gpurun_out/c4_full_${TAG} python tools/step_traffic.py run c4 \
    > gpurun_out/ncu_c4_full_${TAG}.log 2>&1
echo "full capture c4 rc=$?"
# the reports are 2-3 MB per kernel: summarise on the box, bring back the CSVs (gpurun_out/ is
# capped at 64 MiB) and only a small report for the source-level view
python tools/ncu_summary.py full gpurun_out/conv_full_${TAG}.ncu-rep gpurun_out/conv_full_${TAG}_summary.csv > /dev/null
python tools/ncu_summary.py full gpurun_out/c4_full_${TAG}.ncu-rep gpurun_out/c4_full_${TAG}_summary.csv > /dev/null
rm -f gpurun_out/c4_full_${TAG}.ncu-rep gpurun_out/conv_full_${TAG}.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" \
    --profile-from-start off -s 8 -c 6 -o gpurun_out/conv_src_${TAG} python tools/step_traffic.py run c1 \
    > /dev/null 2>&1
echo "source capture rc=$?"
