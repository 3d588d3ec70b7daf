#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/tc_accuracy.py > gpurun_out/r2_tc_quick.json 2> gpurun_out/r2_tc_quick.err; tail -1 gpurun_out/r2_tc_quick.json; tail -2 gpurun_out/r2_tc_quick.err
timeout 300 python -m pytest tests/test_gpu_contraction.py -m gpu -q -s -k "tcgen05 or distillation" > gpurun_out/r2_tc_quick_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "tcgen05 c64|passed|failed|Error" gpurun_out/r2_tc_quick_pytest.log | tail -8
