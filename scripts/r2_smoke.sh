#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r2_smoke.log)"

