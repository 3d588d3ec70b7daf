#!/bin/bash
# Round 2, 8 GPUs: bench line (weak headline, config-4 strong block, reshuffles with checks),
# cross-rank correctness + stress, redistribution variants.
N=${N:-8}
mkdir -p gpurun_out
scripts/with_timeout.sh 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'check',d['result_check'],[k for k,v in d['checks'].items() if not v])
print('e2e',d['e2e'])
print('strong4',{k:v for k,v in d['strong_config4'].items() if k!='config'})
print({k:(round(v['ms'],3), round(v['GB/s']/d['n_gpus'])) for k,v in d['reshuffle'].items() if isinstance(v,dict) and 'ms' in v})
"
tail -3 gpurun_out/r2_bench_n$N.err
scripts/with_timeout.sh 200 env SBB_CHUNK_BYTES=256 OMP_NUM_THREADS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py --backend nccl --cases 20 --stress 500 > gpurun_out/r2_dist_check_n$N.log 2>&1; echo "dist rc=$?"; grep -E "DIST_CHECK|differs|failures|stress" gpurun_out/r2_dist_check_n$N.log | tail -10
for cfg in "SBB_NONE=1" "SBB_P2P_PACK_GRID=148" "SBB_P2P_AUX_GRID=148"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  scripts/with_timeout.sh 150 env SBB_ONLY_REDIST=1 $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/perf_redist.py > gpurun_out/r2_perf_redist_n${N}_$tag.json 2> gpurun_out/r2_perf_redist_n${N}_$tag.err; echo "$cfg rc=$?"; tail -1 gpurun_out/r2_perf_redist_n${N}_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v['ms'],3), round(v['GB/s_per_gpu'])) for k,v in d.items() if isinstance(v,dict) and 'ms' in v})"
done
