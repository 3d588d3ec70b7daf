#!/bin/bash
# Round 2, 1 GPU: new tests (masked one-pass, SB_DEBUG self-check, full-size c64), bench line, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_copy.py tests/test_gpu_dropin.py tests/test_gpu_full_size.py tests/test_gpu_loopback.py -m gpu -x -q -k "mask or sb_debug or complex_float or loopback or dropin_program" > gpurun_out/r2_pytest_c.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_c.log)"
grep -E "Error|assert " gpurun_out/r2_pytest_c.log | head
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n1_c.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'check',d['result_check'],[k for k,v in d['checks'].items() if not v])
print('c64',d.get('contraction_c64'))
print('config1',d['reshuffle'].get('config1'))
print({k:(round(v['ms'],3), round(v['GB/s']/d['n_gpus']), v.get('kernel_launches_per_step')) for k,v in d['reshuffle'].items() if isinstance(v,dict) and 'ms' in v})
"
tail -3 gpurun_out/r2_bench_n1_c.err
python scripts/prof_contract.py c64 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_tc python scripts/prof_contract.py c64 > gpurun_out/r2_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
python scripts/prof_copy.py masked > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:permute_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_masked python scripts/prof_copy.py masked > gpurun_out/r2_ncu_masked.log 2>&1; echo "ncu masked rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
