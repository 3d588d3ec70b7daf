#!/bin/bash
# Round 2, 1 GPU: the whole GPU test-suite, the reference's dist.cpp shapes, the full bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_gpu.log)"
timeout 300 tests/cxx/ref_dist_wrapper --dim='16 16 16 32 16' --reps=5 > gpurun_out/r2_ref_dist.log 2>&1; echo "dist.cpp rc=$?"
sed -n '/>>> GPU tests/,$p' gpurun_out/r2_ref_dist.log | grep -A1 "results for m,n,k,batch_size" | grep -v "^--" | paste - - | awk '{print $6, $7, $8, $9, $10, $11}'
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n1_b.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'check',d['result_check'])
print('e2e',d['e2e'])
print('cpu',d['cpu_baseline'])
print('strong4',{k:v for k,v in d['strong_config4'].items() if k!='config'})
print('c64',d.get('contraction_c64'))
print('refgpu',d.get('reference_gpu'))
print({k:(round(v['ms'],3), round(v['GB/s']/d['n_gpus'])) for k,v in d['reshuffle'].items() if isinstance(v,dict) and 'ms' in v})
"
tail -5 gpurun_out/r2_bench_n1_b.err
