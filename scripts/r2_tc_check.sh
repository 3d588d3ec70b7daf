#!/bin/bash
# Round 2: first run of the tcgen05 complex-float contraction on a B200 (parity, then timing)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_contraction.py -m gpu -x -q -s -k "tcgen05 or distillation" > gpurun_out/r2_tc_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "tcgen05 c64|passed|failed|Error|error" gpurun_out/r2_tc_pytest.log | tail -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n1_a.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'check',d['result_check'],d['checks'])
print('strong4',d.get('strong_config4'))
print('c64',d.get('contraction_c64'))
print({k:(round(v['ms'],3), round(v['GB/s']/d['n_gpus'])) for k,v in d['reshuffle'].items() if isinstance(v,dict) and 'ms' in v})
"
tail -5 gpurun_out/r2_bench_n1_a.err
