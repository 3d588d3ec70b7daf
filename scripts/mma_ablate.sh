#!/bin/bash
# Ablations of contract_mma_kernel (SBB_MMA_DEBUG: 1 no global loads, 2 no barrier, 3 both)
for d in "$@"; do
  SBB_MMA_DEBUG=$d python bench.py --steps 10 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print('debug', os.environ.get('SBB_MMA_DEBUG'), round(d['roofline']['achieved'],2), round(d['roofline']['frac'],4))"
done
