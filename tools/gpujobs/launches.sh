mkdir -p gpurun_out
for cfg in "humanoid humanoid_8192 0" "ant ant_1m 131072"; do set -- $cfg
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_$1.csv python bench.py --workload $2 --envs $3 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>&1
  echo $1 launches: $(grep -c "gpu__time_duration" gpurun_out/r02_launches_$1.csv) step_kernel: $(grep -c step_kernel gpurun_out/r02_launches_$1.csv)
done
