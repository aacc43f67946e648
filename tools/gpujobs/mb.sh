#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
./tools/microbench/syncwarp_cost | tee gpurun_out/r02_syncwarp_cost.txt
