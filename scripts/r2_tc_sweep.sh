#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_tc_sweep.jsonl
for P in 0 8 4 2 1; do
  SBB_TC_PROMOTE=$P timeout 120 python scripts/tc_accuracy.py >> gpurun_out/r2_tc_sweep.jsonl 2>> gpurun_out/r2_tc_sweep.err || echo "P=$P failed rc=$?"
done
SBB_TC_PROMOTE=0 timeout 120 python scripts/tc_accuracy.py exact >> gpurun_out/r2_tc_sweep.jsonl 2>> gpurun_out/r2_tc_sweep.err
SBB_TC_PROMOTE=2 SBB_TC_KSPLIT=37 timeout 120 python scripts/tc_accuracy.py >> gpurun_out/r2_tc_sweep.jsonl 2>> gpurun_out/r2_tc_sweep.err
SBB_TC_PROMOTE=1 SBB_TC_KSPLIT=37 timeout 120 python scripts/tc_accuracy.py >> gpurun_out/r2_tc_sweep.jsonl 2>> gpurun_out/r2_tc_sweep.err
cat gpurun_out/r2_tc_sweep.jsonl; tail -3 gpurun_out/r2_tc_sweep.err
timeout 300 python -m pytest tests/test_gpu_contraction.py -m gpu -q -s -k "tcgen05 or distillation" > gpurun_out/r2_tc_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "tcgen05 c64|passed|failed|Error" gpurun_out/r2_tc_pytest.log | tail -20
timeout 200 oracle/_ref/ref_gpu_bench --reps=5 > gpurun_out/r2_ref_gpu.json 2> gpurun_out/r2_ref_gpu.err; echo "ref gpu rc=$?"; cat gpurun_out/r2_ref_gpu.json; tail -3 gpurun_out/r2_ref_gpu.err
