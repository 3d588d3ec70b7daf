N=${N:-8}
scripts/with_timeout.sh 200 env SBB_CHUNK_BYTES=256 OMP_NUM_THREADS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py --backend nccl --cases 12 > gpurun_out/dist_check_n$N.log 2>&1; echo "dist rc=$?"; grep -E "DIST_CHECK|differs" gpurun_out/dist_check_n$N.log | tail -5
for cfg in "SBB_P2P_SIGNAL=1" "SBB_P2P_SIGNAL=0"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  scripts/with_timeout.sh 150 env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/perf_redist.py > gpurun_out/perf_redist_n${N}_$tag.json 2> gpurun_out/perf_redist_n${N}_$tag.err; echo "$cfg rc=$?"; tail -1 gpurun_out/perf_redist_n${N}_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v['ms'],3), round(v['GB/s_per_gpu'])) for k,v in d.items() if isinstance(v,dict) and 'ms' in v})"
done
