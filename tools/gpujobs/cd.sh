mkdir -p gpurun_out
BXG_LIB=brax_b200/libbxg_cdims.so python -m pytest tests/test_gpu_bitexact.py -q -k "ant or humanoid" > gpurun_out/cd_tests.log 2>&1; tail -3 gpurun_out/cd_tests.log
for rep in 1 2; do for name in main cdims; do
  lib=brax_b200/libbxg_$name.so; [ $name = main ] && lib=brax_b200/libbxg.so
  for wl in humanoid_8192 ant_1m humanoid_512k; do
    BXG_LIB=$lib python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/cd.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['workload'], round(d['value']))"
  done
done; done
tail -3 gpurun_out/cd.err
