"""System <-> JSON.  Compiled model constants travel as small JSON files so the
package works where neither MuJoCo nor the MJCF sources are available."""
from __future__ import annotations

import dataclasses
import json
from typing import Any

import numpy as np

from brax_b200 import base

_CLASSES = {c.__name__: c for c in (
    base.Transform, base.Motion, base.Inertia, base.Link, base.DoF,
    base.Actuator, base.Option, base.System)}


def _enc(x: Any):
  if dataclasses.is_dataclass(x):
    return {'__class__': type(x).__name__,
            **{f.name: _enc(getattr(x, f.name)) for f in dataclasses.fields(x)}}
  if isinstance(x, np.ndarray):
    data = x.astype(np.float64).tolist() if x.dtype.kind == 'f' else x.tolist()
    return {'__ndarray__': data, 'dtype': str(x.dtype), 'shape': list(x.shape)}
  if isinstance(x, (np.floating,)):
    return {'__ndarray__': float(x), 'dtype': str(x.dtype), 'shape': []}
  if isinstance(x, (np.integer,)):
    return int(x)
  if isinstance(x, tuple):
    return {'__tuple__': [_enc(v) for v in x]}
  if isinstance(x, list):
    return [_enc(v) for v in x]
  return x


def _fix_inf(v):
  if isinstance(v, list):
    return [_fix_inf(t) for t in v]
  return v


def _dec(x: Any):
  if isinstance(x, dict):
    if '__class__' in x:
      cls = _CLASSES[x['__class__']]
      return cls(**{k: _dec(v) for k, v in x.items() if k != '__class__'})
    if '__ndarray__' in x:
      return np.array(x['__ndarray__'], dtype=x['dtype']).reshape(x['shape'])
    if '__tuple__' in x:
      return tuple(_dec(v) for v in x['__tuple__'])
  if isinstance(x, list):
    return [_dec(v) for v in x]
  return x


def dumps(sys: base.System) -> str:
  return json.dumps(_enc(sys), indent=1)


def loads(s: str) -> base.System:
  return _dec(json.loads(s))


def save(sys: base.System, path) -> None:
  with open(path, 'w') as f:
    f.write(dumps(sys))


def load(path) -> base.System:
  with open(path, 'r') as f:
    return loads(f.read())
