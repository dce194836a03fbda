#!/usr/bin/env bash
# Round-2 records for profiles/: bench lines of every workload, the reference arm, the ncu launch list of the default bench
# command and one full ncu capture of the dominant kernels.  Usage (1 B200): tools/r2_records.sh
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py > $O/r2_bench_infer.json 2> $O/r2_bench_infer.err || exit 1
python bench.py --workload train > $O/r2_bench_train.json 2>> $O/r2_bench_infer.err
python bench.py --workload router --no-cpu-baseline > $O/r2_bench_router.json 2>> $O/r2_bench_infer.err
python bench.py --workload sweep192 --no-cpu-baseline > $O/r2_bench_sweep192.json 2>> $O/r2_bench_infer.err
python bench.py --workload sweep288 --no-cpu-baseline > $O/r2_bench_sweep288.json 2>> $O/r2_bench_infer.err
python bench.py --graph --steps 50 --no-cpu-baseline --no-train-record > $O/r2_bench_infer_graph.json 2>> $O/r2_bench_infer.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/r2_bench_reference.json 2>> $O/r2_bench_infer.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_bench_infer_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cv_fwd_lean|head_fwd_x3r|head_bwd_x3w|cv_bwd_v4|cv_bwd_staged" -s 40 -c 8 -o $O/r2_bench_kernels \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O/*.ncu-rep
