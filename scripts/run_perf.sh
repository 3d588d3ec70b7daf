python -m pytest tests/test_gpu_copy.py -m gpu -x -q > gpurun_out/pytest_copy.log 2>&1; tail -3 gpurun_out/pytest_copy.log
python scripts/perf_copy.py > gpurun_out/perf_copy_v3.json 2> gpurun_out/perf_copy_v3.err; python -c "
import json; d=json.load(open('gpurun_out/perf_copy_v3.json')); print({k:v['GB/s'] for k,v in d.items()})"
SBB_EPT4=8 SBB_EPT8=16 python scripts/perf_copy.py > gpurun_out/perf_copy_v3b.json 2> gpurun_out/perf_copy_v3b.err; python -c "
import json; d=json.load(open('gpurun_out/perf_copy_v3b.json')); print('EPT4=8,EPT8=16', {k:v['GB/s'] for k,v in d.items()})"
