"""Per-function instruction / stall-sample shares from an ncu report.

  python tools/ncu_breakdown.py gpurun_out/prof.ncu-rep [top_lines]

Reads the source page (needs -lineinfo and --import-source on at capture time)
and attributes every source line of brax_b200/csrc/bxg_core.cuh to the enclosing
BXG_HD function.  Used to write the summaries under profiles/."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
  rep = sys.argv[1]
  top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
  out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, cur, per_line = None, None, {}
  for r in rows:
    if len(r) >= 2 and r[0] in ('File Path', 'File Name'):
      cur = r[1].split('/')[-1]
    elif len(r) > 2 and r[0] == 'Line No':
      hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':
      d = dict(zip(hdr, r))
      key = (cur, int(r[0]))
      inst, samp = int(d['Instructions Executed'] or 0), int(d['# Samples'] or 0)
      a, b, _ = per_line.get(key, (0, 0, ''))
      per_line[key] = (a + inst, b + samp, r[1])
  # function starts from the source of the captured build: pass the git revision as 3rd
  # argument when the tree has moved on since the capture
  if len(sys.argv) > 3:
    text = subprocess.run(['git', 'show', f'{sys.argv[3]}:brax_b200/csrc/bxg_core.cuh'], capture_output=True, text=True, cwd=ROOT).stdout
  else:
    text = open(os.path.join(ROOT, 'brax_b200', 'csrc', 'bxg_core.cuh')).read()
  funcs = []
  for i, l in enumerate(text.split('\n'), 1):
    m = re.match(r'BXG_HD(?:_NOINLINE)? [\w\s\*:<>]*?(\w+)\(', l)
    if m:
      funcs.append((i, m.group(1)))

  def func_of(line):
    name = '?'
    for s, n in funcs:
      if s <= line:
        name = n
    return name
  tot = sum(v[0] for v in per_line.values()) or 1
  tots = sum(v[1] for v in per_line.values()) or 1
  agg, aggs = collections.Counter(), collections.Counter()
  for (f, ln), (inst, samp, _) in per_line.items():
    k = func_of(ln) if f == 'bxg_core.cuh' else f
    agg[k] += inst; aggs[k] += samp
  print(f'total warp instructions {tot}  stall samples {tots}')
  print(f'{"function":30s} {"inst %":>8s} {"samples %":>10s}')
  for k, v in agg.most_common(25):
    print(f'{k:30s} {100 * v / tot:8.2f} {100 * aggs[k] / tots:10.2f}')
  print('--- hottest source lines (by stall samples)')
  for (f, ln), (inst, samp, txt) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{f}:{ln:<5d} inst {100 * inst / tot:5.2f}%  samples {100 * samp / tots:5.2f}%  {txt.strip()[:90]}')


if __name__ == '__main__':
  main()
