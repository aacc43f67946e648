# usage: bash tools/gpujobs/ab_ant.sh lib1 lib2 ...   Ant 1M only, two rounds
mkdir -p gpurun_out
for rep in 1 2; do for name in "$@"; do
  lib=brax_b200/libbxg_$name.so; [ $name = main ] && lib=brax_b200/libbxg.so
  BXG_LIB=$lib python bench.py --workload ant_1m --steps 10 --no-extra --no-cpu-baseline 2>> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['config']['workload'], round(d['value']), d['config']['launch']['envs_per_cta'])"
done; done
