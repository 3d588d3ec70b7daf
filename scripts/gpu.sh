#!/bin/bash
# Build (the loader refuses a library that does not match the sources), then run a command on a B200
# box:  scripts/gpu.sh [--gpus N] [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -m superbblas_b200.build >/dev/null
python -c "import superbblas_b200._lib as l; l.lib()" 
exec /usr/local/graft/bin/gpurun "$@"
