# usage: bash tools/gpujobs/ab_hum.sh lib1 lib2 ...   Humanoid 8192 and 512k, two rounds
mkdir -p gpurun_out
for rep in 1 2; do for name in "$@"; do
  lib=brax_b200/libbxg_$name.so; [ $name = main ] && lib=brax_b200/libbxg.so
  for wl in humanoid_8192 humanoid_512k; do
  BXG_LIB=$lib python bench.py --workload $wl --steps 6 --no-extra --no-cpu-baseline 2>> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['workload'], round(d['value']), d['config']['launch']['envs_per_cta'])"
  done
done; done
