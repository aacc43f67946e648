# usage: bash tools/gpujobs/multi.sh N  -> bench.py on N GPUs (torchrun), both arms
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu_ref.json 2>> gpurun_out/r02_bench_${N}gpu.err
tail -5 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.startswith('{')][-1]
print('main', d['config']['workload'], d['n_gpus'], round(d['value']), 'e2e', round(d['e2e']['value']))
for o in d.get('other_workloads', []):
    print(o.get('workload'), o.get('minv',''), round(o.get('value', 0)), o.get('error',''))
r=[json.loads(l) for l in open('gpurun_out/r02_bench_${N}gpu_ref.json') if l.startswith('{')][-1]; print('reference', round(r['value']), r['cpu_baseline']['cores'])
PY
