python scripts/prof_copy.py perm64 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:permute_kernel -s 1 -c 1 -f -o gpurun_out/prof_perm64_r1d python scripts/prof_copy.py perm64 > gpurun_out/ncu_p64.log 2>&1; echo "ncu perm64 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:permute_kernel -s 1 -c 1 -f -o gpurun_out/prof_perm128_r1d python scripts/prof_copy.py perm128 > gpurun_out/ncu_p128.log 2>&1; echo "ncu perm128 rc=$?"
python scripts/prof_contract.py c64 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contract_mma -s 1 -c 1 -f -o gpurun_out/prof_contract_c64_r1d python scripts/prof_contract.py c64 > gpurun_out/ncu_c64.log 2>&1; echo "ncu contract c64 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
