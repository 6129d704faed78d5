#!/bin/bash
# Scaling sweep on one box with 8 GPUs: bench.py at N = 1, 2, 4, 8 for the default workload (C2, one image per GPU),
# the HD batch (C3, 64 frames per GPU) and the MCU-row sharded giant image (C5, strong scaling).
# usage: tools/scale_run.sh <out.jsonl> [max_gpus]
out=${1:-gpurun_out/scale.jsonl}
maxn=${2:-8}
: > "$out"
port=29600
for wl in "c2" "c3 --batch 64" "c5"; do
  for n in 1 2 4 8; do
    [ "$n" -gt "$maxn" ] && continue
    port=$((port + 1))
    if [ "$n" -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --workload $wl --steps 20 --warmup 5 --no-cpu-baseline >> "$out" 2>> "$out.err"
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --workload $wl --steps 20 --warmup 5 --no-cpu-baseline >> "$out" 2>> "$out.err"
    fi
    echo "workload=$wl n=$n rc=$?" >> "$out.err"
  done
done
grep -c metric "$out"
