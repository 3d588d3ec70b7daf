python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_r1c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r1c.log
python bench.py > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_d.json 2> gpurun_out/bench_ref_d.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_r1d.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench_d.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:permute_kernel -s 3 -c 1 -f -o gpurun_out/prof_perm64_r1d python scripts/prof_copy.py perm64 > gpurun_out/ncu_p64.log 2>&1; echo "ncu perm64 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:permute_kernel -s 3 -c 1 -f -o gpurun_out/prof_perm128_r1d python scripts/prof_copy.py perm > gpurun_out/ncu_p128.log 2>&1; echo "ncu perm128 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:contract_mma -s 1 -c 1 -f -o gpurun_out/prof_contract_c64_r1d python scripts/prof_copy.py contract64 > gpurun_out/ncu_c64.log 2>&1; echo "ncu contract64 rc=$?"
