#!/bin/bash
# domain-randomisation kernels: GPU tests, then the regression checks that the default kernels did not move
cd /root/repo
python -m pytest tests/test_gpu_domain_randomization.py -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-ppo > gpurun_out/dr_bench.json 2> gpurun_out/dr_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/dr_bench.json').read().strip().splitlines()[-1])
print('humanoid_8192', d['value'], 'e2e', d['e2e']['value'])
for v in d.get('other_workloads',[]): print(v.get('workload', v.get('config')), v.get('value'))
P
python tools/dr_bench.py 2>&1 | tail -8
