#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_envs.py -x -q 2>&1 | tail -3
for m in pusher humanoidstandup halfcheetah walker2d hopper; do python tools/model_bench.py $m 65536 2>/dev/null | tail -1; done
