python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_pytest_gpu.log; tail -4 gpurun_out/r02g_pytest_gpu.log
for wl in humanoid_8192 ant_1m humanoid_512k; do python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/r02g.err | tee -a gpurun_out/r02g_bench.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value']), d['config']['launch']['envs_per_cta'])"; done
python tools/ppo_bench.py 12000000 > gpurun_out/r02g_ppo.json 2>> gpurun_out/r02g.err; tail -c 600 gpurun_out/r02g_ppo.json
tail -3 gpurun_out/r02g.err
