#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
bash tools/gpujobs/prof.sh r02_final_ant ant_1m 131072
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_ant.csv python bench.py --workload ant_1m --envs 131072 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>&1
echo launches: $(grep -c "gpu__time_duration" gpurun_out/r02_launches_ant.csv) step_kernel: $(grep -c step_kernel gpurun_out/r02_launches_ant.csv)
ncu -i gpurun_out/prof_r02_final_ant.ncu-rep --page raw --csv > gpurun_out/prof_r02_final_ant_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r02_final_ant.ncu-rep --page source --csv > gpurun_out/prof_r02_final_ant_src.csv 2>/dev/null
ls -la gpurun_out | grep final_ant
