python -m pytest tests/test_gpu_bitexact.py -q -k "lean" > gpurun_out/r02d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_tests.log; tail -15 gpurun_out/r02d_tests.log
for m in ant humanoid humanoid_falls; do BXG_LIB=brax_b200/libbxg_timers.so python tools/phase_timers.py $m > gpurun_out/r02d_phases_$m.json 2>> gpurun_out/r02d.err; cat gpurun_out/r02d_phases_$m.json; done
python bench.py --workload humanoid_8192 --steps 10 --no-cpu-baseline > gpurun_out/r02d_bench.json 2>> gpurun_out/r02d.err
tail -5 gpurun_out/r02d.err
