"""Static SASS instruction count per source function of a kernel variant.
  python tools/static_code_size.py build/obj/bxg_inst_v1.o step_kernel"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj, kern = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
src = open(os.path.join(ROOT, 'brax_b200', 'csrc', 'bxg_core.cuh')).read().split('\n')
funcs = [(i, m.group(1)) for i, l in enumerate(src, 1) for m in [re.match(r'BXG_HD [\w\s\*:<>]*?(\w+)\(', l)] if m]
def func_of(line):
  name = '?'
  for s, n in funcs:
    if s <= line: name = n
  return name
cnt, cur, active = collections.Counter(), None, False
for l in out.split('\n'):
  if l.startswith('.text.'):
    active = kern in l
  m = re.search(r'//## File "([^"]+)", line (\d+)', l)
  if m:
    f, ln = os.path.basename(m.group(1)), int(m.group(2))
    cur = func_of(ln) if f == 'bxg_core.cuh' else f
  elif active and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur:
    cnt[cur] += 1
tot = sum(cnt.values())
print(f'{kern}: {tot} SASS instructions ({tot * 16 / 1024:.0f} KB)')
for k, v in cnt.most_common(24):
  print(f'  {k:32s} {v:6d}  {100 * v / tot:5.1f}%')
