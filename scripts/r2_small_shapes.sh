#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_contraction.py tests/test_gpu_dropin.py -m gpu -x -q -k "not reference_own_contraction" > gpurun_out/r2_pytest_f.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_f.log)"
timeout 300 tests/cxx/ref_dist_wrapper --dim='16 16 16 32 16' --reps=5 > gpurun_out/r2_ref_dist2.log 2>&1; echo "dist.cpp rc=$?"
sed -n '/>>> GPU tests/,$p' gpurun_out/r2_ref_dist2.log | grep -A1 "results for m,n,k,batch_size" | grep -v "^--" | paste - - | sed 's/results for m,n,k,batch_size: //; s/Time in contracting nn//'
