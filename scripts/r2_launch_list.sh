#!/bin/bash
# ncu launch list of the bench command (per-launch device times: cold-cache, serialised -- compare SHARES)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r2_ncu_launches.log 2>&1; echo "rc=$?"
grep -c "contract_mma\|contract_reduce" gpurun_out/r2_bench_launches.csv
