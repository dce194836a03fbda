#!/bin/bash
# 1 -> 8 GPU weak-scaling run of bench.py (one node).  Usage: tools/scale_run.sh [workload] [outfile]
W=${1:-infer}; OUT=${2:-gpurun_out/scale_$W.jsonl}; : > $OUT
python bench.py --gpus 1 --workload $W --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null >> $OUT
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --workload $W --steps 100 --warmup 5 2>/dev/null >> $OUT
done
python - "$OUT" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
base = rows[0]["value"]
for r in rows:
    print(f'N={r["n_gpus"]} pairs/s={r["value"]:.0f} ms/step={r["ms_per_step"]:.4f} scaling_eff={r["value"]/(base*r["n_gpus"]):.3f} e2e={r["e2e"]["value"]:.0f} cv_frac={r["roofline"]["frac"]}')
PY
