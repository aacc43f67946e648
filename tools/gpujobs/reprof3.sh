#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
bash tools/gpujobs/prof.sh r02_final_humanoid humanoid_8192
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_humanoid.csv python bench.py --workload humanoid_8192 --envs 0 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>&1
echo launches: $(grep -c "gpu__time_duration" gpurun_out/r02_launches_humanoid.csv) step_kernel: $(grep -c step_kernel gpurun_out/r02_launches_humanoid.csv)
