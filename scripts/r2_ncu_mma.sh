#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_contract.py c128 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract_mma_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_mma2 python scripts/prof_contract.py c128 > gpurun_out/r2_ncu_mma2.log 2>&1; echo "ncu mma rc=$?"
