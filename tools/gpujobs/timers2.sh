#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for m in ant humanoid; do
  BXG_LIB=brax_b200/libbxg_timers.so python tools/phase_timers.py $m > gpurun_out/r02k_phases_$m.json 2>&1; cat gpurun_out/r02k_phases_$m.json
done
