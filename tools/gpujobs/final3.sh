#!/bin/bash
# round-2 closing run after the per-env-model kernels: full GPU suite, sanitizer, default bench line, smoke
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck.txt; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.txt; tail -4 gpurun_out/r02_sanitizer_racecheck.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/dr_bench.py 2>&1 | tail -3
python bench.py > gpurun_out/r02_final3_bench_line.json 2> gpurun_out/r02_final3_bench.err; tail -c 600 gpurun_out/r02_final3_bench_line.json
