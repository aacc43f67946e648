mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_pytest_gpu.log; tail -4 gpurun_out/r02_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/gpujobs/final1.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_default_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ppo > /dev/null 2>&1; grep -c step_kernel gpurun_out/r02_launches_default_bench.csv
