mkdir -p gpurun_out
T0=$(date +%s); python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "default bench wall $(( $(date +%s) - T0 )) s"
T0=$(date +%s); python bench.py --impl reference > gpurun_out/r02_final_bench_ref.json 2>> gpurun_out/r02_final_bench.err; echo "reference arm wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_final_bench.json') if l.startswith('{')][-1]
print('main', d['config']['workload'], round(d['value']), 'e2e', round(d['e2e']['value']), 'fp32 frac', round(d['fp32']['frac'],3), 'launch', d['config']['launch'], 'clocks', d['clocks'])
for o in d.get('other_workloads', []):
    print(o.get('workload'), o.get('minv',''), o.get('state_io','')[:4], round(o.get('value', 0)), o.get('error',''))
print('cpu_baseline', d.get('cpu_baseline'))
r=[json.loads(l) for l in open('gpurun_out/r02_final_bench_ref.json') if l.startswith('{')][-1]; print('reference', round(r['value']), r['cpu_baseline']['cores'])
PY
