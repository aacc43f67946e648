for m in ant humanoid; do BXG_LIB=brax_b200/libbxg_timers.so python tools/phase_timers.py $m > gpurun_out/r02e_phases_$m.json 2>> gpurun_out/r02e.err; cat gpurun_out/r02e_phases_$m.json; done
for wl in humanoid_8192 ant_1m; do BXG_LIB=brax_b200/libbxg_timers.so python bench.py --workload $wl --steps 5 --no-extra --no-cpu-baseline 2>> gpurun_out/r02e.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('timers build', d['config']['workload'], round(d['value']))"; done
tail -3 gpurun_out/r02e.err
