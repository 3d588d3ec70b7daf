#!/bin/bash
# First GPU call of the next round (1 GPU, ~3 minutes): validate the opt-in contraction kernels that were
# written without GPU access (bodies checked on the CPU by tests/test_row_kernel_emulation.py), then time
# the reference's tests/dist.cpp shapes with and without them.  Outputs under gpurun_out/.
#   gpurun --timeout 600 -- 'bash scripts/validate_optin_kernels.sh'
set -u
mkdir -p gpurun_out
run() { # name, env...
    local name=$1; shift
    env "$@" python -m pytest tests/test_gpu_contraction.py tests/test_gpu_dropin.py -m gpu -x -q \
        > gpurun_out/optin_$name.log 2>&1
    echo "$name: rc=$? $(tail -1 gpurun_out/optin_$name.log)"
}
run baseline SBB_NONE=1
run row SBB_ROW_KERNEL=1
run dot SBB_DOT_KERNEL=1
run simt_order SBB_SIMT_ORDER=1
run all SBB_ROW_KERNEL=1 SBB_DOT_KERNEL=1 SBB_SIMT_ORDER=1
for cfg in "SBB_NONE=1" "SBB_ROW_KERNEL=1 SBB_DOT_KERNEL=1" "SBB_SIMT_ORDER=1"; do
    tag=$(echo "$cfg" | tr ' =' '__')
    env $cfg tests/cxx/ref_dist_wrapper --dim='16 16 16 32 16' --reps=5 > gpurun_out/ref_dist_$tag.log 2>&1
    echo "dist.cpp [$cfg] rc=$?"
    sed -n '/>>> GPU tests/,$p' gpurun_out/ref_dist_$tag.log | grep -A1 "results for m,n,k,batch_size: \(1,1,\|4,4,\|12,12,\|49152,3,3\|49152,16,16\)" | grep -v "^--"
done
