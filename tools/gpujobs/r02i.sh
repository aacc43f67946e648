mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_pytest_gpu.log; tail -4 gpurun_out/r02i_pytest_gpu.log
for wl in humanoid_8192 ant_1m humanoid_512k; do python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/r02i.err | tee -a gpurun_out/r02i_bench.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value']), d['config']['launch'])"; done
for m in ant humanoid; do BXG_LIB=brax_b200/libbxg_timers.so python tools/phase_timers.py $m > gpurun_out/r02i_phases_$m.json 2>> gpurun_out/r02i.err; python -c "
import json; d=json.load(open('gpurun_out/r02i_phases_$m.json')); print('$m', d['share_of_warp_cycles'])"; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
tail -3 gpurun_out/r02i.err
