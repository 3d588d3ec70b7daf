#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loopback.py tests/test_gpu_copy.py -m gpu -x -q -k "halo_gather or storage_staging" > gpurun_out/r2_pytest_e.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_e.log)"; grep -E "Error|assert " gpurun_out/r2_pytest_e.log | head
