"""Shared-memory wavefronts per source line of an ncu report (needs --import-source on):
  python tools/ncu_smem.py gpurun_out/prof.ncu-rep [top]
'excess' = wavefronts beyond the ideal for the access (bank conflicts)."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, ex, wf, src = None, None, collections.Counter(), collections.Counter(), {}
for r in rows:
  if len(r) >= 2 and r[0] in ('File Path', 'File Name'):
    cur = r[1].split('/')[-1]
  elif len(r) > 2 and r[0] == 'Line No':
    hdr = r
  elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':
    d = dict(zip(hdr, r))
    k = (cur, int(r[0]))
    ex[k] += int(d['L1 Wavefronts Shared Excessive'] or 0); wf[k] += int(d['L1 Wavefronts Shared'] or 0); src[k] = r[1].strip()[:100]
T = sum(wf.values()) or 1
print(f'{rep}: shared wavefronts {T}, excessive {sum(ex.values())} ({100 * sum(ex.values()) / T:.1f}%)')
print('-- lines with the most excess wavefronts')
for k, v in ex.most_common(top):
  print(f'  {k[0]}:{k[1]:<5d} excess {100 * v / T:5.2f}%  (line {100 * wf[k] / T:5.2f}%)  {src[k]}')
print('-- lines with the most wavefronts')
for k, v in wf.most_common(top):
  print(f'  {k[0]}:{k[1]:<5d} {100 * v / T:5.2f}%  {src[k]}')
