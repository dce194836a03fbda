#!/usr/bin/env bash
# 1/2/4/8-GPU runs of bench.py on one box (what the driver does at round end), for profiles/.
#   gpurun --gpus 8 -- tools/scale_run.sh [workload]      -> gpurun_out/r2_scale_<workload>_1_2_4_8.jsonl
cd "$(dirname "$0")/.."
W=${1:-infer}
OUT=gpurun_out/r2_scale_${W}_1_2_4_8.jsonl
: > $OUT
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 100 --warmup 5 --workload $W --no-cpu-baseline --no-train-record 2>/dev/null | grep '^{' >> $OUT
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 100 --warmup 5 --workload $W --no-cpu-baseline --no-train-record 2>/dev/null | grep '^{' >> $OUT
  fi
done
python - "$OUT" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1])]
v1, e1 = rows[0]["value"], rows[0]["e2e"]["value"]
for r in rows:
    n = r["n_gpus"]
    print(f"N={n}: {r['value']:.0f} pairs/s (eff {r['value']/(n*v1):.3f}), serial {r['value_serial']:.0f}, e2e {r['e2e']['value']:.0f} (eff {r['e2e']['value']/(n*e1):.3f}, "
          f"{r['e2e']['h2d_GBps_per_rank']} GB/s per rank, {r['e2e']['frac_of_h2d_ceiling']} of the H2D ceiling {r['e2e']['h2d_ceiling_GBps_per_rank']} GB/s)")
PY
