#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python bench.py --workload ant_1m --steps 200 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ant_1m 200 steps', round(d['value']), 'nonfinite', d['nonfinite_envs'])"
python bench.py --workload humanoid_512k --steps 100 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('humanoid_512k 100 steps', round(d['value']), 'nonfinite', d['nonfinite_envs'])"
python tests/rollout_report.py ant 1024 1000 > gpurun_out/r02_rollout_ant_1024x1000.json 2>> gpurun_out/soak.err
python tests/rollout_report.py humanoid 512 1000 > gpurun_out/r02_rollout_humanoid_512x1000.json 2>> gpurun_out/soak.err
python - <<'P'
import json
for m in ('ant_1024','humanoid_512'):
    d=json.load(open(f'gpurun_out/r02_rollout_{m}x1000.json')); print(m, {k: d[k] for k in list(d)[:8] if not isinstance(d[k], (list, dict))})
P
tail -2 gpurun_out/soak.err
