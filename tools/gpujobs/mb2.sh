#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
./tools/microbench/lds128_patterns | tee gpurun_out/r02_lds128_patterns.txt
