#!/bin/bash
# After adding the SB_TRACK_TIME reports (host code only): the new GPU test, the core GPU tests and smoke().
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_report.py tests/test_capi.py::test_reports_of_the_public_calls -x -q \
    > gpurun_out/r2_pytest_report.log 2>&1; echo "report rc=$? $(tail -1 gpurun_out/r2_pytest_report.log)"
timeout 50 python -m pytest tests/test_gpu_copy.py tests/test_gpu_contraction.py tests/test_gpu_loopback.py -m gpu -x -q \
    > gpurun_out/r2_pytest_core.log 2>&1; echo "core rc=$? $(tail -1 gpurun_out/r2_pytest_core.log)"
SB_TRACK_TIME=1 timeout 25 python -c "
import __graft_entry__ as g, superbblas_b200 as sb
g.smoke(); print(sb.reportTimings())" > gpurun_out/r2_smoke_tracked.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r2_smoke_tracked.log
