#!/bin/bash
N=${N:-4}
mkdir -p gpurun_out
for cfg in "SBB_NONE=1" "SBB_P2P_PACK_GRID=148"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  scripts/with_timeout.sh 150 env SBB_ONLY_REDIST=1 $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/perf_redist.py > gpurun_out/r2_perf_redist_n${N}_$tag.json 2> gpurun_out/r2_perf_redist_n${N}_$tag.err; echo "$cfg rc=$?"; tail -1 gpurun_out/r2_perf_redist_n${N}_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v['ms'],3), round(v['GB/s_per_gpu'])) for k,v in d.items() if isinstance(v,dict) and 'ms' in v})"
done
