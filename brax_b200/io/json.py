"""Trajectory export in the reference viewer's JSON schema.

Restates `brax/io/json.py:96-156`: a dict with the system's `dt` and link names,
`geoms` grouped by link name (name, link_idx, pos, rot, rgba, size) and
`states.x` = one {pos, rot} per frame, so a rollout of the B200 step can be
loaded into the reference's three.js visualiser (`brax/visualizer/`) for an
eyeball parity check.
"""
from __future__ import annotations

import json
from typing import List

import numpy as np

_GEOM_TYPE_NAMES = {0: 'Plane', 1: 'HeightMap', 2: 'Sphere', 3: 'Capsule', 5: 'Cylinder', 6: 'Box', 7: 'Mesh'}
_DEFAULT_RGBA = [0.4, 0.33, 0.26, 1.0]


def _list(a):
  a = a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)
  return a.astype(np.float64).tolist()


def dumps(sys, states: List) -> str:
  """sys: a brax_b200 System; states: pipeline states of ONE env (x.pos [L,3], x.rot [L,4])."""
  for s in states:
    if (len(s.x.pos.shape), len(s.x.rot.shape)) != (2, 2):
      raise RuntimeError(
          'Expected state.x position and rotation to have 2 shape dimensions but '
          f'received len(pos.shape)={len(s.x.pos.shape)} and len(rot.shape)={len(s.x.rot.shape)}')
  link_names = [n or f'link {i}' for i, n in enumerate(sys.link_names)] + ['world']
  geoms = {}
  for g in range(len(sys.geom_type)):
    link_idx = int(sys.geom_bodyid[g]) - 1
    geoms.setdefault(link_names[link_idx], []).append({
        'name': _GEOM_TYPE_NAMES[int(sys.geom_type[g])], 'link_idx': link_idx,
        'pos': _list(sys.geom_pos[g]), 'rot': _list(sys.geom_quat[g]),
        'rgba': _DEFAULT_RGBA, 'size': _list(sys.geom_size[g])})
  d = {
      'name': 'System', 'link_names': list(sys.link_names),
      'opt': {'timestep': float(sys.opt.timestep), 'name': 'Option'},
      'geoms': geoms,
      'states': {'x': [{'pos': _list(s.x.pos), 'rot': _list(s.x.rot), 'name': 'Transform'} for s in states]},
  }
  return json.dumps(d)


def save(path: str, sys, states: List) -> None:
  with open(path, 'w') as f:
    f.write(dumps(sys, states))
