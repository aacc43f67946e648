python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_reference_big.py -q > gpurun_out/r02f_bitexact.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_bitexact.log; tail -4 gpurun_out/r02f_bitexact.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_bitexact.py --deselect tests/test_gpu_reference_big.py > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_pytest_gpu.log; tail -4 gpurun_out/r02f_pytest_gpu.log
for wl in humanoid_8192 ant_1m humanoid_512k; do python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/r02f.err | tee -a gpurun_out/r02f_bench.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value']), d['config']['launch'])"; done
tail -3 gpurun_out/r02f.err
