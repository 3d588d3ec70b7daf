#!/bin/bash
# The NCCL send/recv transport (SBB_P2P=0; also the automatic fallback when the arenas cannot be mapped)
mkdir -p gpurun_out
scripts/with_timeout.sh 200 env SBB_P2P=0 SBB_CHUNK_BYTES=256 OMP_NUM_THREADS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py --backend nccl --cases 20 --stress 200 > gpurun_out/r2_dist_check_n2_nccl.log 2>&1; echo "dist rc=$?"; grep -E "DIST_CHECK|differs|failures|stress" gpurun_out/r2_dist_check_n2_nccl.log | tail -5
