mkdir -p gpurun_out
for name in "$@"; do
  lib=brax_b200/libbxg_$name.so; [ $name = main ] && lib=brax_b200/libbxg.so
  for fv in 1 4; do for wl in humanoid_8192 humanoid_512k; do
    BXG_FORCE_VARIANT=$fv BXG_LIB=$lib python bench.py --workload $wl --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/v4.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name variant $fv', d['config']['workload'], round(d['value']), d['config']['launch'])"
  done; done
done
tail -3 gpurun_out/v4.err
