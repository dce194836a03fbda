#!/usr/bin/env bash
# H2D ceiling sweep on one box: 1/2/4/8 ranks, bound and unbound.  Usage (8-GPU box): tools/h2d_sweep.sh > gpurun_out/h2d.jsonl
cd "$(dirname "$0")/.."
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r2_topo.txt 2>&1
for n in 1 2 4 8; do
  for flag in "" "--no-bind"; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) tools/h2d_ceiling.py $flag 2>/dev/null | grep '^{'
  done
done
