#!/bin/bash
# usage: tools/run_checks.sh [VAR=val]stage ...   (each stage in its own process, 170 s limit)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for item in "$@"; do
  st="${item##*:}"; envs=""
  if [[ "$item" == *:* ]]; then envs="${item%:*}"; fi
  echo "######## $item" >> gpurun_out/check.log
  env $envs timeout 170 python tools/gpu_check.py $st >> gpurun_out/check.log 2>&1
  echo "exit=$?" >> gpurun_out/check.log
done
tail -n 120 gpurun_out/check.log
