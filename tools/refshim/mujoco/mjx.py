"""mjx containers and the two colliders the path needs.

NOT the reference: `collision` restates mjx/_src/collision_primitive.py `plane_sphere`
and `plane_capsule` and the pair bookkeeping of collision_driver.py (static pair list,
friction = max, solref / solimp mixed with equal solmix) -- the same statement
oracle/bxg_oracle.c and brax_b200/base.py `contact_pairs` make.  The static pair table is
handed over by the generator script (PAIRS)."""
import dataclasses

import numpy as np

from flax import struct
from jax import numpy as jp

PAIRS = None   # brax_b200.base.ContactPairs of the model being run


class PyTreeNode:
  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    c = dataclasses.dataclass(frozen=True, kw_only=True)(cls)
    c.replace = lambda self, **upd: dataclasses.replace(self, **upd)
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


def collision(sys, d):
  cp = PAIRS
  n = len(cp.geom1)
  dt = np.asarray(d.geom_xpos).dtype
  dist, pos, frame = np.zeros(n, dt), np.zeros((n, 3), dt), np.zeros((n, 3, 3), dt)
  for k in range(n):
    g1, g2 = int(cp.geom1[k]), int(cp.geom2[k])
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
