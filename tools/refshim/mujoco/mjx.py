"""mjx containers and the three colliders the path needs.

NOT the reference: `collision` restates mjx/_src/collision_primitive.py `plane_sphere`,
`plane_capsule` and `capsule_capsule` (with math.closest_segment_to_segment_points and
math.make_frame) and the pair bookkeeping of collision_driver.py (static pair list,
friction = max, solref / solimp mixed with equal solmix) -- the same statement
oracle/bxg_oracle.c and brax_b200/base.py `contact_pairs` make.  The static pair table is
handed over by the generator script (PAIRS)."""
import dataclasses

import numpy as np

from flax import struct
from jax import numpy as jp

PAIRS = None   # brax_b200.base.ContactPairs of the model being run


def _tree_replace_all(self, params):
  """mjx's PyTreeNode.tree_replace ({'a.b.c': value}: nested `replace`), which brax.System inherits."""
  def rep(node, attr, val):
    if len(attr) == 1:
      return node.replace(**{attr[0]: val})
    return node.replace(**{attr[0]: rep(getattr(node, attr[0]), attr[1:], val)})
  new = self
  for k, v in params.items():
    new = rep(new, k.split('.'), v)
  return new


class PyTreeNode:
  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    c = dataclasses.dataclass(frozen=True, kw_only=True)(cls)
    c.replace = lambda self, **upd: dataclasses.replace(self, **upd)
    c.tree_replace = _tree_replace_all
    struct.register_dataclass(c)


class Option(PyTreeNode):
  timestep: float = None


class Model(PyTreeNode):
  nq: int = struct.field(pytree_node=False, default=0)
  nv: int = struct.field(pytree_node=False, default=0)
  nu: int = struct.field(pytree_node=False, default=0)
  opt: Option = None
  geom_bodyid: np.ndarray = struct.field(pytree_node=False, default=None)
  geom_pos: np.ndarray = None
  geom_quat: np.ndarray = None


class Contact(PyTreeNode):
  dist: np.ndarray = None
  pos: np.ndarray = None
  frame: np.ndarray = None
  includemargin: np.ndarray = None
  friction: np.ndarray = None
  solref: np.ndarray = None
  solreffriction: np.ndarray = None
  solimp: np.ndarray = None
  dim: np.ndarray = struct.field(pytree_node=False, default=None)
  geom1: np.ndarray = None
  geom2: np.ndarray = None
  geom: np.ndarray = None
  efc_address: np.ndarray = struct.field(pytree_node=False, default=None)


class Data(PyTreeNode):
  ncon: int = struct.field(pytree_node=False, default=0)
  geom_xpos: np.ndarray = None
  geom_xmat: np.ndarray = None
  contact: Contact = None


def make_data(sys):
  return Data(ncon=0 if PAIRS is None else len(PAIRS.geom1))


def _normalize_with_norm(x):
  is_zero = np.allclose(x, 0.0)
  n = 0.0 if is_zero else np.linalg.norm(x)
  return x / (n + 1e-6 * (n == 0.0)), n


def _closest_segment_point(a, b, pt):
  ab = b - a
  t = np.dot(pt - a, ab) / (np.dot(ab, ab) + 1e-6)
  return a + np.clip(t, 0.0, 1.0) * ab


def _closest_segment_to_segment_points(a0, a1, b0, b1):
  dir_a, len_a = _normalize_with_norm(a1 - a0)
  dir_b, len_b = _normalize_with_norm(b1 - b0)
  half_a, half_b = len_a * 0.5, len_b * 0.5
  a_mid, b_mid = a0 + dir_a * half_a, b0 + dir_b * half_b
  trans = a_mid - b_mid
  dab, dat, dbt = dir_a.dot(dir_b), dir_a.dot(trans), dir_b.dot(trans)
  denom = 1 - dab * dab
  orig_t_a = (-dat + dab * dbt) / (denom + 1e-6)
  orig_t_b = dbt + orig_t_a * dab
  t_a, t_b = np.clip(orig_t_a, -half_a, half_a), np.clip(orig_t_b, -half_b, half_b)
  best_a, best_b = a_mid + dir_a * t_a, b_mid + dir_b * t_b
  new_a = _closest_segment_point(a0, a1, best_b)
  new_b = _closest_segment_point(b0, b1, best_a)
  d1 = (new_a - best_b).dot(new_a - best_b)
  d2 = (best_a - new_b).dot(best_a - new_b)
  return (new_a, best_b) if d1 < d2 else (best_a, new_b)


def _make_frame(a):
  a, _ = _normalize_with_norm(a)
  b = np.array([0.0, 1.0, 0.0]) if -0.5 < a[1] < 0.5 else np.array([0.0, 0.0, 1.0])
  b = b - a * a.dot(b)
  b, _ = _normalize_with_norm(b)
  return np.stack([a, b, np.cross(a, b)])


def _capsule_capsule(d, cp, k):
  g1, g2 = int(cp.geom1[k]), int(cp.geom2[k])
  c1, c2 = np.asarray(d.geom_xpos[g1]), np.asarray(d.geom_xpos[g2])
  seg1 = np.asarray(d.geom_xmat[g1])[:, 2] * float(cp.a_half[k])
  seg2 = np.asarray(d.geom_xmat[g2])[:, 2] * float(cp.half_len[k])
  pt1, pt2 = _closest_segment_to_segment_points(c1 - seg1, c1 + seg1, c2 - seg2, c2 + seg2)
  n, dist = _normalize_with_norm(pt2 - pt1)      # _sphere_sphere
  if dist == 0.0:
    n = np.array([1.0, 0.0, 0.0])
  r1, r2 = float(cp.a_radius[k]), float(cp.radius[k])
  dist = dist - (r1 + r2)
  return dist, pt1 + n * (r1 + dist * 0.5), _make_frame(n)


def collision(sys, d):
  cp = PAIRS
  n = len(cp.geom1)
  dt = np.asarray(d.geom_xpos).dtype
  dist, pos, frame = np.zeros(n, dt), np.zeros((n, 3), dt), np.zeros((n, 3, 3), dt)
  for k in range(n):
    g1, g2 = int(cp.geom1[k]), int(cp.geom2[k])
    if int(cp.kind[k]) == 2:
      dist[k], pos[k], frame[k] = _capsule_capsule(d, cp, k)
      continue
    nrm = np.asarray(d.geom_xmat[g1])[:, 2]
    ppos, c = np.asarray(d.geom_xpos[g1]), np.asarray(d.geom_xpos[g2])
    r = dt.type(cp.radius[k])
    if int(cp.kind[k]) == 0:
      fr = np.asarray(cp.frame[k], dt)          # math.make_frame(n): constant for a fixed plane
    else:
      axis = np.asarray(d.geom_xmat[g2])[:, 2]
      b, bn = _normalize_with_norm(axis - nrm * nrm.dot(axis))
      if bn < 0.5:
        b = np.array([0.0, 1.0, 0.0], dt) if -0.5 < nrm[1] < 0.5 else np.array([0.0, 0.0, 1.0], dt)
      fr = np.stack([nrm, b, np.cross(nrm, b)])
      c = c + axis * dt.type(cp.half_len[k])
    dist[k] = np.dot(c - ppos, nrm) - r
    pos[k] = c - nrm * (r + 0.5 * dist[k])
    frame[k] = fr
  mu = np.asarray(cp.friction, dt)
  con = Contact(
      dist=jp.array(dist), pos=jp.array(pos), frame=jp.array(frame), includemargin=jp.zeros(n),
      friction=jp.array(np.stack([mu, mu, np.zeros(n, dt), np.zeros(n, dt), np.zeros(n, dt)], 1)),
      solref=jp.array(np.asarray(cp.solref, dt)), solreffriction=jp.zeros((n, 2)), solimp=jp.array(np.asarray(cp.solimp, dt)),
      dim=np.full(n, 3), geom1=jp.array(np.asarray(cp.geom1)), geom2=jp.array(np.asarray(cp.geom2)),
      geom=jp.array(np.stack([np.asarray(cp.geom1), np.asarray(cp.geom2)], 1)), efc_address=np.arange(n))
  return d.replace(contact=con)
