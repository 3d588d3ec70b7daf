#!/bin/bash
# ncu captures (one kernel each, after the same command ran without ncu)
mkdir -p gpurun_out
python scripts/prof_contract.py c128 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract_mma_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_mma python scripts/prof_contract.py c128 > gpurun_out/r2_ncu_mma.log 2>&1; echo "ncu mma rc=$?"
python scripts/prof_copy.py masked > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:permute_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_masked4 python scripts/prof_copy.py masked > gpurun_out/r2_ncu_masked4.log 2>&1; echo "ncu masked rc=$?"
python scripts/prof_contract.py c64 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_tc18 python scripts/prof_contract.py c64 > gpurun_out/r2_ncu_tc18.log 2>&1; echo "ncu tc rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
