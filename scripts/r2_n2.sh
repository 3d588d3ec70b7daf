#!/bin/bash
# Round 2, 2 GPUs: correctness of the cross-rank path, bench with result checks, peer-store probe.
N=${N:-2}
mkdir -p gpurun_out
scripts/with_timeout.sh 200 env SBB_CHUNK_BYTES=256 OMP_NUM_THREADS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py --backend nccl --cases 20 --stress 1000 > gpurun_out/r2_dist_check_n$N.log 2>&1; echo "dist rc=$?"; grep -E "DIST_CHECK|differs|failures|stress" gpurun_out/r2_dist_check_n$N.log | tail -5
scripts/with_timeout.sh 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'check',d['result_check'],d['checks'])
print('strong4',d.get('strong_config4'))
print({k:(round(v['ms'],3), round(v['GB/s']/d['n_gpus'])) for k,v in d['reshuffle'].items() if isinstance(v,dict) and 'ms' in v})
"
tail -5 gpurun_out/r2_bench_n$N.err
timeout 120 python scripts/tc_accuracy.py > gpurun_out/r2_tc_default.json 2>&1; cat gpurun_out/r2_tc_default.json | tail -1
