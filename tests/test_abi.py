"""The C-ABI library loads and exports every symbol include/bxg.h declares.
No compute calls here (CPU box)."""
import ctypes
import os
import re
import subprocess

import pytest

from tests.conftest import ROOT


def _build():
  import __graft_entry__ as g
  g.build()


def test_header_symbols_exported():
  _build()
  hdr = open(os.path.join(ROOT, 'include', 'bxg.h')).read()
  hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
  declared = set(re.findall(r'\b(bxg_[a-z_]+)\s*\(', hdr))
  assert {'bxg_init', 'bxg_step', 'bxg_model_create', 'bxg_model_destroy', 'bxg_last_error'} <= declared
  out = subprocess.run(['nm', '-D', '--defined-only', os.path.join(ROOT, 'brax_b200', 'libbxg.so')],
                       capture_output=True, text=True, check=True).stdout
  exported = set(re.findall(r'\bT (bxg_[a-z_]+)', out))
  assert declared <= exported, declared - exported


def test_loads_and_fails_loudly_without_cuda(ant):
  _build()
  from brax_b200 import native
  lib = native.lib()
  assert lib.bxg_abi_version() == 4
  import torch
  if torch.cuda.is_available():
    pytest.skip('CUDA present: the no-device error path is exercised on CPU boxes')
  desc, keep = native.make_desc(ant)
  h = ctypes.c_void_p()
  rc = lib.bxg_model_create(ctypes.byref(desc), 0, ctypes.byref(h))
  assert rc == 2 and not h.value            # BXG_E_CUDA, no handle: no CPU fallback
  assert b'no CUDA device' in lib.bxg_last_error()
  with pytest.raises(RuntimeError):
    native.NativeModel(ant, 0)


def test_desc_struct_matches_header_layout(ant, humanoid):
  """ctypes mirror of BxgModelDesc: field order/count must follow bxg.h."""
  from brax_b200 import native
  hdr = open(os.path.join(ROOT, 'include', 'bxg.h')).read()
  head = 'typedef struct BxgModelDesc {'
  body = hdr[hdr.index(head) + len(head):hdr.index('} BxgModelDesc;')]
  body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
  names = []
  for decl in body.split(';'):
    decl = decl.strip()
    if not decl or decl.startswith('typedef'):
      continue
    for part in decl.split(','):
      names.append(re.findall(r'([A-Za-z_0-9]+)\s*(?:\[\d+\])?$', part.strip())[0])
  assert names == [f[0] for f in native.ModelDesc._fields_]
  st = hdr[hdr.index('typedef struct BxgState {'):hdr.index('} BxgState;')]
  st = re.sub(r'/\*.*?\*/', '', st, flags=re.S)
  assert tuple(re.findall(r'float\*\s+([a-z_]+);', st)) == native.STATE_FIELDS
  for s in (ant, humanoid):
    d, _ = native.make_desc(s)
    assert d.nv == s.nv and d.ncon == len(s.contact_pairs().geom1)


def test_only_test_infrastructure_touches_the_oracle():
  """oracle/ is test infrastructure: nothing under brax_b200/ or tools/ may import it (only tests/,
  __graft_entry__.smoke() / build() and bench.py's CPU legs do)."""
  import re
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  pat = re.compile(r'^\s*(from\s+oracle\b|import\s+oracle\b)', re.M)
  offenders = []
  for top in ('brax_b200', 'tools'):
    for dirpath, _, files in os.walk(os.path.join(root, top)):
      for f in files:
        if f.endswith(('.py', '.cu', '.cuh', '.h')):
          path = os.path.join(dirpath, f)
          text = open(path, errors='ignore').read()
          includes_oracle = any('oracle' in l for l in text.split('\n') if l.lstrip().startswith('#include'))
          if (f.endswith('.py') and pat.search(text)) or includes_oracle:
            offenders.append(os.path.relpath(path, root))
  assert not offenders, offenders
