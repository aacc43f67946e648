"""A/B sweep of launch / variant knobs on the GPU box (one bench.py run per configuration):
  python tools/variant_sweep.py gpurun_out/sweep.json [only-substring]
Each entry: label, workload, environment overrides."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def _lib(name):
  return {'BXG_LIB': os.path.join(ROOT, 'brax_b200', f'libbxg_{name}.so')}


CONFIGS = [   # edit per experiment; alternative builds come from tools/build_alt.py (_lib('name'))
    ('hum_default', 'humanoid_8192', {}),
    ('hum512k_default', 'humanoid_512k', {}),
    ('ant_default', 'ant_1m', {}),
]
out_path = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else ''
res = []
for label, wl, env in CONFIGS:
  if only and only not in label:
    continue
  e = dict(os.environ); e.update(env)
  p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', wl, '--steps', '10', '--warmup', '3',
                      '--no-cpu-baseline', '--no-extra'], env=e, capture_output=True, text=True, timeout=150)
  line = None
  for l in p.stdout.splitlines():
    if l.startswith('{'):
      line = json.loads(l)
  r = {'label': label, 'workload': wl, 'env': {k: os.path.basename(v) for k, v in env.items()}}
  if line:
    r.update(value=line['value'], ms_per_step=line['ms_per_step'], nonfinite=line.get('nonfinite_envs'),
             clocks=line.get('clocks', {}).get('sm_mhz'), launch=line.get('config', {}).get('launch'))
  else:
    r['error'] = (p.stderr or p.stdout)[-800:]
  print(json.dumps(r), flush=True)
  res.append(r)
  with open(out_path, 'w') as f:
    json.dump(res, f, indent=1)
