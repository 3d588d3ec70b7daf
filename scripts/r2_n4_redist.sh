#!/bin/bash
# 4 GPUs: correctness of the paced exchange (many rounds, random partitions) and the redistribution
N=${N:-4}
mkdir -p gpurun_out
scripts/with_timeout.sh 200 env SBB_CHUNK_BYTES=256 OMP_NUM_THREADS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py --backend nccl --cases 20 --stress 300 > gpurun_out/r2_dist_check_n$N.log 2>&1; echo "dist rc=$?"; grep -E "DIST_CHECK|differs|failures|stress" gpurun_out/r2_dist_check_n$N.log | tail -6
scripts/with_timeout.sh 150 env SBB_ONLY_REDIST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/perf_redist.py > gpurun_out/r2_perf_redist_n${N}_paced.json 2> gpurun_out/r2_perf_redist_n${N}_paced.err; echo "redist rc=$?"; tail -1 gpurun_out/r2_perf_redist_n${N}_paced.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v['ms'],3), round(v['GB/s_per_gpu'])) for k,v in d.items() if isinstance(v,dict) and 'ms' in v})"
