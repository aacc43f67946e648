mkdir -p gpurun_out
python -m pytest tests/test_gpu_ppo_kernels.py tests/test_gpu_ppo.py -q > gpurun_out/r02j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_tests.log; tail -12 gpurun_out/r02j_tests.log
python tools/ppo_bench.py 12000000 > gpurun_out/r02j_ppo.json 2>> gpurun_out/r02j.err; tail -c 700 gpurun_out/r02j_ppo.json
tail -3 gpurun_out/r02j.err
