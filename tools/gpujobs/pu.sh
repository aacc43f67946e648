#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_domain_randomization.py tests/test_gpu_envs.py -x -q 2>&1 | tail -3
bash tools/gpujobs/ab.sh main
python bench.py --workload humanoid_512k --steps 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('humanoid_512k', round(d['value']))"
python tools/dr_bench.py | tail -2
