#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_clean.json 2> gpurun_out/r2_bench_n1_clean.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n1_clean.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'check',d['result_check'],d['clocks'])
print('strong4',d['strong_config4']['value'],d['strong_config4']['ms_per_step'])
print('c64',d['contraction_c64']['TFLOP/s'],d['contraction_c64']['kernel_ms'],d['contraction_c64']['clocks_under_this_kernel'])
"
