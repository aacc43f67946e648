python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_reference_big.py -q > gpurun_out/r02c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_tests.log; tail -5 gpurun_out/r02c_tests.log
bash tools/gpujobs/prof.sh r02c_ant ant_1m 131072
bash tools/gpujobs/prof.sh r02c_humanoid humanoid_8192
