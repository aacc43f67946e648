# usage: bash tools/gpujobs/prof.sh <tag> <workload> [envs]
# one full ncu capture (source-level) of the step kernel + a launch list of a short bench run
TAG=$1; WL=$2; ENVS=${3:-0}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG} -f \
  python bench.py --workload $WL --envs $ENVS --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/prof_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --workload $WL --envs $ENVS --steps 3 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/prof_${TAG}.log
