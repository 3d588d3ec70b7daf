#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_tc_atmem.jsonl
SBB_TC_ATMEM=1 timeout 120 python scripts/tc_accuracy.py >> gpurun_out/r2_tc_atmem.jsonl 2>> gpurun_out/r2_tc_atmem.err || echo "atmem failed rc=$?"
timeout 120 python scripts/tc_accuracy.py >> gpurun_out/r2_tc_atmem.jsonl 2>> gpurun_out/r2_tc_atmem.err
cat gpurun_out/r2_tc_atmem.jsonl; tail -3 gpurun_out/r2_tc_atmem.err
SBB_TC_ATMEM=1 timeout 300 python -m pytest tests/test_gpu_contraction.py -m gpu -q -s -k "tcgen05 or distillation" > gpurun_out/r2_tc_atmem_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "tcgen05 c64|passed|failed|Error" gpurun_out/r2_tc_atmem_pytest.log | tail -12
