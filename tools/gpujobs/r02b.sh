mkdir -p gpurun_out
python -m pytest tests/test_gpu_bitexact.py -q > gpurun_out/r02b_bitexact.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_bitexact.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_bitexact.py > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_pytest_gpu.log
for wl in humanoid_8192 ant_1m humanoid_512k; do python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline >> gpurun_out/r02b_bench.jsonl 2>> gpurun_out/r02b_bench.err; done
tail -30 gpurun_out/r02b_bitexact.log; tail -5 gpurun_out/r02b_pytest_gpu.log
python -c "
import json
for l in open('gpurun_out/r02b_bench.jsonl'): d=json.loads(l); print(d['config']['workload'], round(d['value']))
"
