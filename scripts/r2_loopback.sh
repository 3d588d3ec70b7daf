#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loopback.py tests/test_gpu_copy.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2_loopback.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_loopback.log)"
grep -E "Error|error|assert" gpurun_out/r2_loopback.log | head -20
