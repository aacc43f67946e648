#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
bash tools/gpujobs/prof.sh r02_final_ant ant_1m 131072
bash tools/gpujobs/prof.sh r02_final_humanoid humanoid_8192
bash tools/gpujobs/launches.sh
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ls -la gpurun_out | grep "final_\|launches"
