#!/bin/bash
# Multi-GPU measurements on one box (run under `gpurun --gpus 8`): C1 at 1/2/4/8 ranks, then C3,
# C4, C2 and C5 at 8.  One JSON line per run, appended to gpurun_out/scale_${TAG}.jsonl.
TAG=${1:-r2}
OUT=gpurun_out/scale_${TAG}.jsonl
mkdir -p gpurun_out
: > $OUT
run() {  # N, extra args...
  local N=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --no-extras --no-cpu-baseline "$@" >> $OUT 2>> gpurun_out/scale_${TAG}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + N)) bench.py --gpus $N --no-cpu-baseline "$@" >> $OUT 2>> gpurun_out/scale_${TAG}.err
  fi
  echo "N=$N $* rc=$?"
}
for N in 1 2 4 8; do run $N --steps 50 --warmup 5; done
run 8 --config c3 --steps 30 --warmup 5
run 1 --config c4 --steps 6 --warmup 3 --batch 64
run 8 --config c4 --steps 6 --warmup 3 --batch 64
run 8 --config c2 --steps 20
run 8 --config c5 --steps 30
