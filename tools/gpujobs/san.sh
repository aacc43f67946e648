mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck.txt; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.txt; tail -4 gpurun_out/r02_sanitizer_racecheck.txt
python -m pytest tests/test_gpu_envs.py -q -k episode_metrics 2>&1 | tail -2
