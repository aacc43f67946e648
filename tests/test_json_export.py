"""Trajectory export schema (reference brax/io/json.py:96-156)."""
import json
import types

import numpy as np
import pytest


def test_schema_and_shapes(ant):
  from brax_b200.io import json as bjson
  from oracle import oracle as O
  o = O.Oracle(ant)
  st = o.init(np.asarray(ant.init_q, np.float32)[None], np.zeros((1, ant.nv), np.float32))
  frames = []
  for _ in range(3):
    o.step(st, np.zeros((1, ant.nu), np.float32), 5)
    frames.append(types.SimpleNamespace(x=types.SimpleNamespace(pos=st['x_pos'][0].copy(), rot=st['x_rot'][0].copy())))
  d = json.loads(bjson.dumps(ant, frames))
  assert d['opt']['timestep'] == pytest.approx(0.01)
  assert len(d['states']['x']) == 3
  assert np.array(d['states']['x'][0]['pos']).shape == (9, 3) and np.array(d['states']['x'][0]['rot']).shape == (9, 4)
  assert 'world' in d['geoms'] and d['geoms']['world'][0]['name'] == 'Plane' and d['geoms']['world'][0]['link_idx'] == -1
  assert {g['name'] for g in d['geoms']['torso']} == {'Sphere', 'Capsule'}
  assert 'link 2' in d['geoms']           # unnamed links get 'link i' (json.py:117)
  bad = [types.SimpleNamespace(x=types.SimpleNamespace(pos=st['x_pos'], rot=st['x_rot']))]   # batched: rejected
  with pytest.raises(RuntimeError):
    bjson.dumps(ant, bad)


@pytest.mark.parametrize('name', ['pusher', 'swimmer', 'humanoidstandup', 'reacher', 'inverted_pendulum',
                                  'inverted_double_pendulum', 'hopper', 'walker2d', 'halfcheetah', 'humanoid'])
def test_every_shipped_model_exports(name):
  from brax_b200 import envs_assets
  from brax_b200.io import json as bjson
  s = envs_assets.load(name)
  n = s.num_links()
  frame = types.SimpleNamespace(x=types.SimpleNamespace(pos=np.zeros((n, 3), np.float32),
                                                        rot=np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1))))
  d = json.loads(bjson.dumps(s, [frame]))
  assert len(d['states']['x']) == 1 and sum(len(g) for g in d['geoms'].values()) == len(s.geom_type)
