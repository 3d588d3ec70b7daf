#!/bin/bash
# contract_mma_kernel variants: SBB_MMA_BK=8 (64x64x8, 4 stages, 2 CTA/SM), 16 (64x64x16, 2 stages, 2 CTA/SM),
# 17 (64x64x16, 3 stages, 1 CTA/SM); SBB_MMA_BN=32 (64x32x8, 4 stages, 3 CTA/SM)
for v in "$@"; do
  export $v
  timeout 300 python -m pytest tests/test_gpu_contraction.py -x -q 2>&1 | tail -1
  python bench.py --steps 10 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print('$v', round(d['roofline']['achieved'],2), round(d['roofline']['frac'],4))"
  unset ${v%%=*}
done
