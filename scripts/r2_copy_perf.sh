#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/perf_copy.py > gpurun_out/r2_perf_copy.json 2> gpurun_out/r2_perf_copy.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_perf_copy.json'))
print({k:v['GB/s'] for k,v in d.items()})"
timeout 300 python -m pytest tests/test_gpu_copy.py -m gpu -x -q > gpurun_out/r2_pytest_g.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_g.log)"
