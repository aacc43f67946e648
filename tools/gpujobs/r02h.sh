mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for cfg in "humanoid_8192 0 full" "humanoid_8192 0 lean" "ant_1m 131072 full" "ant_1m 131072 lean" "humanoid_512k 65536 full"; do
  set -- $cfg; F=""; [ $3 = lean ] && F="--lean"
  ncu --metrics $M --clock-control none -k regex:step_kernel -s 4 -c 3 --csv --log-file gpurun_out/r02h_dram_$1_$3.csv \
    python bench.py --workload $1 --envs $2 $F --steps 3 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>> gpurun_out/r02h.err
done
bash tools/gpujobs/prof.sh r02_final_ant ant_1m 131072
bash tools/gpujobs/prof.sh r02_final_humanoid humanoid_8192
python tests/rollout_report.py ant 1024 1000 > gpurun_out/r02_rollout_ant_1024x1000.json 2>> gpurun_out/r02h.err
python tests/rollout_report.py humanoid 512 1000 > gpurun_out/r02_rollout_humanoid_512x1000.json 2>> gpurun_out/r02h.err
tail -3 gpurun_out/r02h.err; ls -la gpurun_out | tail -12
