mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02a_bench_ref.json 2>> gpurun_out/r02a_bench.err
for lib in libbxg.so libbxg_nofmad.so; do for wl in humanoid_8192 ant_1m; do BXG_LIB=brax_b200/$lib python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline >> gpurun_out/r02a_nofmad.jsonl 2>> gpurun_out/r02a_bench.err; done; done
tail -3 gpurun_out/r02a_pytest_gpu.log; python -c "
import json
for l in open('gpurun_out/r02a_nofmad.jsonl'): d=json.loads(l); print(d['config']['workload'], round(d['value']))
"
