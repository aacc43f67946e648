"""Builds an alternative libbxg with extra -D flags for A/B experiments on the GPU box:
  python tools/build_alt.py t576 -DBXG_G32_MAXT=576     ->  brax_b200/libbxg_t576.so
select it with BXG_LIB=brax_b200/libbxg_t576.so (brax_b200/native.py)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
name, defs = sys.argv[1], sys.argv[2:]
csrc = os.path.join(ROOT, 'brax_b200', 'csrc')
objdir = os.path.join(ROOT, 'build', 'alt_' + name)
os.makedirs(objdir, exist_ok=True)
nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
flags = [f for f in g.NVCC_FLAGS if f not in ('-Xptxas', '-v')] + defs
procs, objs = [], []
for v in range(g.N_VARIANTS):
  o = os.path.join(objdir, f'v{v}.o'); objs.append(o)
  procs.append(subprocess.Popen([nvcc] + flags + [f'-DBXG_VARIANT={v}', '-c', os.path.join(csrc, 'bxg_inst.cu'), '-o', o]))
for unit in ('api', 'train'):
  o = os.path.join(objdir, f'{unit}.o'); objs.append(o)
  procs.append(subprocess.Popen([nvcc] + flags + ['-c', os.path.join(csrc, f'bxg_{unit}.cu'), '-o', o]))
assert all(p.wait() == 0 for p in procs)
out = os.path.join(ROOT, 'brax_b200', f'libbxg_{name}.so')
subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '--shared', '-Xcompiler', '-fPIC'] + objs + ['-o', out], check=True)
print(out)
