#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_quick.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'peak',d['roofline']['peak'],'check',d['result_check'], d['rel_err_vs_cublas'])
"
tail -2 gpurun_out/r2_bench_quick.err
timeout 600 python -m pytest tests/test_gpu_contraction.py -m gpu -x -q > gpurun_out/r2_pytest_quick.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_quick.log)"
