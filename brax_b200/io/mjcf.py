"""MJCF -> System compiler that does not need the MuJoCo C library.

The reference builds its `System` by handing the XML to MuJoCo's model compiler
and copying compiled fields (reference `brax/io/mjcf.py:317-489`,
`mujoco.MjModel.from_xml_string` at `:509`).  MuJoCo is not available in this
build environment, so this module restates the parts of that compiler that the
`generalized` hot path consumes:

  * body fusing of joint-less bodies            (reference `io/mjcf.py:76-98`)
  * defaults / compiler / option / custom tags  (MJCF schema)
  * inertia-from-geom for sphere / capsule / box / cylinder / ellipsoid
  * `dof_invweight0`, `body_invweight0`         (MuJoCo `engine_setconst.c`, recalled)
  * joint / actuator / geom fields copied by `load_model` (`io/mjcf.py:331-416`)
  * collision pair enumeration (contype/conaffinity, parent-child exclusion)

Supported model class: kinematic trees of free / hinge / slide joints with
primitive geoms; colliding pairs plane-sphere (the only pair Ant and Humanoid
have).  Anything else raises NotImplementedError rather
than silently producing different physics.

All arithmetic here is float64 NumPy; `System` casts to float32 the way the
reference does with `jax.tree.map(jp.array, sys)` (`io/mjcf.py:478`).
"""
from __future__ import annotations

import copy
import math as _math
from typing import Dict, List, Optional, Tuple
from xml.etree import ElementTree

import numpy as np

from brax_b200 import base

# MuJoCo geom type ids (mjtGeom)
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE = 0, 1, 2, 3
GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 4, 5, 6, 7
_GEOM_TYPES = {
    'plane': GEOM_PLANE, 'sphere': GEOM_SPHERE, 'capsule': GEOM_CAPSULE,
    'ellipsoid': GEOM_ELLIPSOID, 'cylinder': GEOM_CYLINDER, 'box': GEOM_BOX,
}
_MJ_MINVAL = 1e-15

_DEFAULT_SOLREF = (0.02, 1.0)
_DEFAULT_SOLIMP = (0.9, 0.95, 0.001, 0.5, 2.0)


# ----------------------------------------------------------------------------
# small float64 quaternion helpers (w, x, y, z)
# ----------------------------------------------------------------------------
def _vec(s: Optional[str], default) -> np.ndarray:
  if s is None:
    return np.array(default, dtype=np.float64)
  return np.array([float(t) for t in s.split()], dtype=np.float64)


def _qmul(u, v):
  return np.array([
      u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3],
      u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
      u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1],
      u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0],
  ])


def _qrot(q, v):
  s, u = q[0], q[1:]
  return 2 * np.dot(u, v) * u + (s * s - np.dot(u, u)) * v + 2 * s * np.cross(u, v)


def _qnorm(q):
  n = np.linalg.norm(q)
  return q / n if n > 0 else np.array([1.0, 0, 0, 0])


def _q2mat(q):
  w, x, y, z = _qnorm(q)
  return np.array([
      [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
  ])


def _mat2q(m):
  """Rotation matrix -> unit quaternion (w>=0 branch preferred)."""
  t = np.trace(m)
  if t > 0:
    s = _math.sqrt(t + 1.0) * 2
    q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s,
                  (m[1, 0] - m[0, 1]) / s])
  elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
    s = _math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
    q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s,
                  (m[0, 2] + m[2, 0]) / s])
  elif m[1, 1] > m[2, 2]:
    s = _math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
    q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s,
                  (m[1, 2] + m[2, 1]) / s])
  else:
    s = _math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
    q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s,
                  (m[1, 2] + m[2, 1]) / s, 0.25 * s])
  return _qnorm(q)


def _orientation(a: Dict[str, str], deg: bool) -> np.ndarray:
  """quat | axisangle | euler (xyz, intrinsic) -> unit quaternion (MJCF orientation attributes)."""
  if 'quat' in a:
    return _qnorm(_vec(a['quat'], [1, 0, 0, 0]))
  if 'axisangle' in a:
    v = _vec(a['axisangle'], [0, 0, 1, 0])
    ang = np.deg2rad(v[3]) if deg else v[3]
    ax = v[:3] / np.linalg.norm(v[:3])
    return np.concatenate([[_math.cos(ang / 2)], ax * _math.sin(ang / 2)])
  if 'euler' in a:
    e = _vec(a['euler'], [0, 0, 0])
    e = np.deg2rad(e) if deg else e
    q = np.array([1.0, 0, 0, 0])
    for k in range(3):   # default eulerseq "xyz", intrinsic rotations
      ax = np.eye(3)[k]
      q = _qmul(q, np.concatenate([[_math.cos(e[k] / 2)], ax * _math.sin(e[k] / 2)]))
    return _qnorm(q)
  if 'xyaxes' in a or 'zaxis' in a:
    raise NotImplementedError('xyaxes / zaxis orientations are not supported')
  return np.array([1.0, 0, 0, 0])


def _z_to(vec):
  """Quaternion rotating +z onto `vec` (MuJoCo mjuu_z2quat convention)."""
  v = vec / np.linalg.norm(vec)
  z = np.array([0.0, 0.0, 1.0])
  axis = np.cross(z, v)
  s = np.linalg.norm(axis)
  if s < 1e-10:
    return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1.0, 0, 0])
  axis /= s
  ang = _math.atan2(s, v[2])
  return np.concatenate([[_math.cos(ang / 2)], axis * _math.sin(ang / 2)])


# ----------------------------------------------------------------------------
# XML pre-processing: defaults and body fusing
# ----------------------------------------------------------------------------
def _fmt6(v) -> np.ndarray:
  """The reference re-serialises offset vectors with '%f' (6 decimals)."""
  return np.array([float('%f' % t) for t in v])


def _reframe(elem, ppos, pquat):
  """Re-express a child element of a dissolved body in the parent's frame
  (behaviour of reference `_offset`, `io/mjcf.py:44-73`)."""
  pos = _vec(elem.attrib.get('pos'), [0, 0, 0])
  quat = _vec(elem.attrib.get('quat'), [1, 0, 0, 0])
  ft = elem.attrib.get('fromto')
  if ft:
    v = _vec(ft, [0] * 6)
    a = ppos + _qrot(pquat, v[0:3])
    b = ppos + _qrot(pquat, v[3:6])
    elem.attrib['fromto'] = ' '.join('%f' % t for t in np.concatenate([a, b]))
    return
  npos = ppos + _qrot(pquat, pos)
  nquat = _qmul(pquat, quat)
  elem.attrib['pos'] = ' '.join('%f' % t for t in npos)
  elem.attrib['quat'] = ' '.join('%f' % t for t in nquat)


def fuse_bodies(elem: ElementTree.Element) -> None:
  """Dissolves joint-less bodies into their parent, depth first.

  Semantics of reference `_fuse_bodies` (`io/mjcf.py:76-98`): children of the
  dissolved body are moved into the parent frame only when the dissolved body's
  `pos` is non-zero (a pure rotation is dropped, exactly as upstream).
  """
  for child in list(elem):
    fuse_bodies(child)
    if child.tag != 'body':
      continue
    if child.find('joint') is not None or child.find('freejoint') is not None:
      continue
    cpos = _vec(child.attrib.get('pos'), [0, 0, 0])
    cquat = _vec(child.attrib.get('quat'), [1, 0, 0, 0])
    for g in list(child):
      if g.tag in ('body', 'geom', 'site', 'camera') and (cpos != 0).any():
        _reframe(g, cpos, cquat)
      elem.append(g)
    elem.remove(child)


class _Defaults:
  """Flat view of <default> classes: class name -> tag -> attribute dict."""

  def __init__(self, root: ElementTree.Element):
    self.classes: Dict[str, Dict[str, Dict[str, str]]] = {'main': {}}
    for d in root.findall('default'):
      self._walk(d, 'main', {})

  def _walk(self, node, name, inherited):
    cur = {k: dict(v) for k, v in inherited.items()}
    for ch in node:
      if ch.tag == 'default':
        continue
      cur.setdefault(ch.tag, {}).update(ch.attrib)
    self.classes[name] = cur
    for ch in node.findall('default'):
      self._walk(ch, ch.attrib['class'], cur)

  def resolve(self, elem, tag, cls):
    out = dict(self.classes.get(cls or 'main', self.classes['main']).get(tag, {}))
    out.update(elem.attrib)
    return out


# ----------------------------------------------------------------------------
# geom mass properties (MuJoCo user_objects.cc mjCGeom::GetVolume / SetInertia)
# ----------------------------------------------------------------------------
def _geom_volume(gtype: int, size: np.ndarray) -> float:
  if gtype == GEOM_SPHERE:
    return 4.0 / 3.0 * _math.pi * size[0] ** 3
  if gtype == GEOM_CAPSULE:
    h = 2 * size[1]
    return _math.pi * (size[0] ** 2 * h + 4.0 / 3.0 * size[0] ** 3)
  if gtype == GEOM_CYLINDER:
    return _math.pi * size[0] ** 2 * 2 * size[1]
  if gtype == GEOM_ELLIPSOID:
    return 4.0 / 3.0 * _math.pi * size[0] * size[1] * size[2]
  if gtype == GEOM_BOX:
    return 8.0 * size[0] * size[1] * size[2]
  return 0.0


def _geom_inertia(gtype: int, size: np.ndarray, mass: float) -> np.ndarray:
  """Diagonal inertia in the geom frame for a solid primitive of given mass."""
  if gtype == GEOM_SPHERE:
    i = 2.0 * mass * size[0] ** 2 / 5.0
    return np.array([i, i, i])
  if gtype == GEOM_CAPSULE:
    r, h = size[0], 2 * size[1]
    sphere_mass = mass * 4 * r / (4 * r + 3 * h)
    cyl_mass = mass - sphere_mass
    ix = cyl_mass * (3 * r * r + h * h) / 12.0
    iz = cyl_mass * r * r / 2.0
    si = 2.0 * sphere_mass * r * r / 5.0
    ix += si + sphere_mass * h * (3 * r + 2 * h) / 8.0
    iz += si
    return np.array([ix, ix, iz])
  if gtype == GEOM_CYLINDER:
    r, h = size[0], 2 * size[1]
    ix = mass * (3 * r * r + h * h) / 12.0
    return np.array([ix, ix, mass * r * r / 2.0])
  if gtype == GEOM_ELLIPSOID:
    a, b, c = size
    return mass / 5.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
  if gtype == GEOM_BOX:
    a, b, c = size
    return mass / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
  return np.zeros(3)


# ----------------------------------------------------------------------------
# intermediate spec
# ----------------------------------------------------------------------------
class _Geom:
  __slots__ = ('name', 'type', 'size', 'pos', 'quat', 'mass', 'inertia',
               'contype', 'conaffinity', 'friction', 'solref', 'solimp',
               'solmix', 'priority', 'condim', 'margin', 'gap', 'body')


class _Joint:
  __slots__ = ('name', 'type', 'pos', 'axis', 'range', 'limited', 'armature',
               'damping', 'stiffness', 'solref', 'solimp', 'ref')


class _Body:
  __slots__ = ('name', 'pos', 'quat', 'parent', 'joints', 'geoms', 'ipos',
               'iquat', 'inertia', 'mass')


def _parse_geom(a: Dict[str, str], density_default=1000.0, deg: bool = True) -> _Geom:
  g = _Geom()
  g.name = a.get('name', '')
  tname = a.get('type', 'sphere')
  if tname not in _GEOM_TYPES:
    raise NotImplementedError(f'geom type "{tname}" is not supported')
  g.type = _GEOM_TYPES[tname]
  size = _vec(a.get('size'), [0, 0, 0])
  size = np.concatenate([size, np.zeros(3 - len(size))]) if len(size) < 3 else size[:3]
  g.pos = _vec(a.get('pos'), [0, 0, 0])
  g.quat = _orientation(a, deg)
  if a.get('fromto'):
    if g.type not in (GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX, GEOM_ELLIPSOID):
      raise NotImplementedError('fromto only for capsule/cylinder/box/ellipsoid')
    ft = _vec(a['fromto'], [0] * 6)
    p0, p1 = ft[0:3], ft[3:6]
    g.pos = 0.5 * (p0 + p1)
    # MuJoCo aligns the geom z axis with (from - to)
    g.quat = _z_to(p0 - p1)
    size = size.copy()
    size[1] = 0.5 * np.linalg.norm(p1 - p0)
  g.size = size
  density = float(a.get('density', density_default))
  if 'mass' in a:
    g.mass = float(a['mass'])
  elif g.type == GEOM_PLANE:
    g.mass = 0.0
  else:
    g.mass = density * _geom_volume(g.type, size)
  g.inertia = _geom_inertia(g.type, size, g.mass)
  g.contype = int(a.get('contype', 1))
  g.conaffinity = int(a.get('conaffinity', 1))
  g.condim = int(a.get('condim', 3))
  g.friction = _vec(a.get('friction'), [1.0, 0.005, 0.0001])
  if len(g.friction) < 3:
    g.friction = np.concatenate(
        [g.friction, np.array([1.0, 0.005, 0.0001])[len(g.friction):]])
  g.solref = _vec(a.get('solref'), _DEFAULT_SOLREF)
  g.solimp = _vec(a.get('solimp'), _DEFAULT_SOLIMP)
  if len(g.solimp) < 5:
    g.solimp = np.concatenate(
        [g.solimp, np.array(_DEFAULT_SOLIMP)[len(g.solimp):]])
  g.solmix = float(a.get('solmix', 1.0))
  g.priority = int(a.get('priority', 0))
  g.margin = float(a.get('margin', 0.0))
  g.gap = float(a.get('gap', 0.0))
  return g


def _parse_joint(a: Dict[str, str], tag: str, deg: bool) -> _Joint:
  j = _Joint()
  j.name = a.get('name', '')
  tname = 'free' if tag == 'freejoint' else a.get('type', 'hinge')
  j.type = {'free': 0, 'ball': 1, 'slide': 2, 'hinge': 3}[tname]
  if j.type == 1:
    raise NotImplementedError('ball joints not supported')
  j.pos = _vec(a.get('pos'), [0, 0, 0])
  axis = _vec(a.get('axis'), [0, 0, 1])
  j.axis = axis / np.linalg.norm(axis)
  rng = _vec(a.get('range'), [0, 0])
  lim = a.get('limited', 'auto')
  if lim == 'auto':
    j.limited = 'range' in a  # compiler autolimits=true (MuJoCo >= 2.3 default)
  else:
    j.limited = lim == 'true'
  if j.type == 3 and deg:
    rng = np.deg2rad(rng)
  j.range = rng
  j.ref = float(a.get('ref', 0.0))
  if j.type == 0:
    j.armature, j.damping, j.stiffness = 0.0, 0.0, 0.0
    if tag != 'freejoint':
      j.armature = float(a.get('armature', 0.0))
      j.damping = float(a.get('damping', 0.0))
      j.stiffness = float(a.get('stiffness', 0.0))
  else:
    j.armature = float(a.get('armature', 0.0))
    j.damping = float(a.get('damping', 0.0))
    j.stiffness = float(a.get('stiffness', 0.0))
  j.solref = _vec(a.get('solreflimit'), _DEFAULT_SOLREF)
  j.solimp = _vec(a.get('solimplimit'), _DEFAULT_SOLIMP)
  if len(j.solimp) < 5:   # a partial solimp keeps MuJoCo's defaults for the rest
    j.solimp = np.concatenate([j.solimp, np.array(_DEFAULT_SOLIMP)[len(j.solimp):]])
  return j


def _body_inertial(b: _Body) -> None:
  """inertiafromgeom: com, principal axes and moments from the body's geoms."""
  mass = sum(g.mass for g in b.geoms)
  if mass < _MJ_MINVAL:
    b.mass, b.ipos = 0.0, np.zeros(3)
    b.iquat, b.inertia = np.array([1.0, 0, 0, 0]), np.zeros(3)
    return
  com = sum(g.mass * g.pos for g in b.geoms) / mass
  tensor = np.zeros((3, 3))
  for g in b.geoms:
    if g.mass <= 0:
      continue
    r = _q2mat(g.quat)
    ig = r @ np.diag(g.inertia) @ r.T
    d = g.pos - com
    tensor += ig + g.mass * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
  # principal axes, eigenvalues in decreasing order (mju_eig3 convention)
  w, v = np.linalg.eigh(0.5 * (tensor + tensor.T))
  order = np.argsort(-w)
  w, v = w[order], v[:, order]
  if np.allclose(w, w[0], rtol=1e-12, atol=1e-300):
    v = np.eye(3)
  else:
    # canonical signs so the result is deterministic, then right-handed
    for k in range(3):
      idx = np.argmax(np.abs(v[:, k]))
      if v[idx, k] < 0:
        v[:, k] = -v[:, k]
    if np.linalg.det(v) < 0:
      v[:, 2] = -v[:, 2]
  b.mass, b.ipos = mass, com
  b.iquat, b.inertia = _mat2q(v), w


# ----------------------------------------------------------------------------
# invweight0 (MuJoCo engine_setconst.c set0, recalled) via explicit Jacobians
# ----------------------------------------------------------------------------
def _qpos0_kinematics(bodies: List[_Body]):
  """World pose of every body at qpos0 (all hinge/slide at 0)."""
  xpos, xquat = [], []
  for b in bodies:
    if b.parent < 0:
      pp, pq = np.zeros(3), np.array([1.0, 0, 0, 0])
    else:
      pp, pq = xpos[b.parent], xquat[b.parent]
    xpos.append(pp + _qrot(pq, b.pos))
    xquat.append(_qnorm(_qmul(pq, b.quat)))
  return xpos, xquat


def _invweight0(bodies: List[_Body], nv: int):
  xpos, xquat = _qpos0_kinematics(bodies)
  # dof table: (body, kind, world axis, world anchor)
  dofs = []
  for bi, b in enumerate(bodies):
    for j in b.joints:
      if j.type == 0:
        for k in range(3):
          dofs.append((bi, 'lin', np.eye(3)[k], None))
        r = _q2mat(xquat[bi])
        for k in range(3):  # free rotational dofs are body-frame axes
          dofs.append((bi, 'ang', r[:, k], xpos[bi]))
      else:
        ax = _qrot(xquat[bi], j.axis)
        anchor = xpos[bi] + _qrot(xquat[bi], j.pos)
        dofs.append((bi, 'lin' if j.type == 2 else 'ang', ax, anchor))
  assert len(dofs) == nv

  def ancestors(bi):
    out = set()
    while bi >= 0:
      out.add(bi)
      bi = bodies[bi].parent
    return out

  def jac(point, bi):
    jp_, jr = np.zeros((3, nv)), np.zeros((3, nv))
    anc = ancestors(bi)
    for d, (db, kind, ax, anchor) in enumerate(dofs):
      if db not in anc:
        continue
      if kind == 'lin':
        jp_[:, d] = ax
      else:
        jr[:, d] = ax
        jp_[:, d] = np.cross(ax, point - anchor)
    return jp_, jr

  m = np.zeros((nv, nv))
  coms, jacs = [], []
  for bi, b in enumerate(bodies):
    c = xpos[bi] + _qrot(xquat[bi], b.ipos)
    coms.append(c)
    jp_, jr = jac(c, bi)
    jacs.append((jp_, jr))
    if b.mass <= 0:
      continue
    r = _q2mat(_qmul(xquat[bi], b.iquat))
    iw = r @ np.diag(b.inertia) @ r.T
    m += b.mass * jp_.T @ jp_ + jr.T @ iw @ jr
  arm = []
  for b in bodies:
    for j in b.joints:
      arm.extend([j.armature] * (6 if j.type == 0 else 1))
  m += np.diag(arm)
  minv = np.linalg.inv(m)

  dof_inv = np.diag(minv).copy()
  d = 0
  for b in bodies:
    for j in b.joints:
      if j.type == 0:
        dof_inv[d:d + 3] = dof_inv[d:d + 3].mean()
        dof_inv[d + 3:d + 6] = dof_inv[d + 3:d + 6].mean()
        d += 6
      else:
        d += 1
  body_inv = np.zeros((len(bodies), 2))
  for bi in range(len(bodies)):
    jp_, jr = jacs[bi]
    body_inv[bi, 0] = np.trace(jp_ @ minv @ jp_.T) / 3.0
    body_inv[bi, 1] = np.trace(jr @ minv @ jr.T) / 3.0
  return dof_inv, body_inv, m


# ----------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------
_CUSTOM_DEFAULTS = {
    'matrix_inv_iterations': 10,
    'solver_maxls': 20,
}


def loads(xml: str) -> base.System:
  """Builds a System from an MJCF string (reference `io/mjcf.py:500-511`)."""
  root = ElementTree.fromstring(xml)
  fuse_bodies(root)
  defaults = _Defaults(root)

  comp = root.find('compiler')
  comp = comp.attrib if comp is not None else {}
  deg = comp.get('angle', 'degree') == 'degree'
  if comp.get('coordinate', 'local') != 'local':
    raise NotImplementedError('only local coordinates are supported')

  opt = root.find('option')
  opt = opt.attrib if opt is not None else {}
  timestep = float(opt.get('timestep', 0.002))
  iterations = int(opt.get('iterations', 100))
  gravity = _vec(opt.get('gravity'), [0, 0, -9.81])
  viscosity = float(opt.get('viscosity', 0.0))
  density = float(opt.get('density', 0.0))
  if opt.get('integrator', 'Euler') not in ('Euler',):
    raise NotImplementedError('Only euler integration is supported.')
  if opt.get('cone', 'pyramidal') != 'pyramidal':
    raise NotImplementedError('Only pyramidal cone friction is supported.')

  custom: Dict[str, np.ndarray] = {}
  for c in root.findall('custom'):
    for n in c.findall('numeric'):
      custom[n.attrib['name']] = _vec(n.attrib.get('data'), [0.0])

  world = root.find('worldbody')
  bodies: List[_Body] = []
  world_geoms: List[_Geom] = []
  for ge in world.findall('geom'):
    g = _parse_geom(defaults.resolve(ge, 'geom', ge.attrib.get('class')), deg=deg)
    g.body = -1
    world_geoms.append(g)

  def walk(elem, parent, childclass):
    for be in elem.findall('body'):
      cc = be.attrib.get('childclass', childclass)
      b = _Body()
      b.name = be.attrib.get('name', '')
      b.pos = _vec(be.attrib.get('pos'), [0, 0, 0])
      b.quat = _orientation(be.attrib, deg)
      b.parent = parent
      b.joints, b.geoms = [], []
      for je in list(be):
        if je.tag in ('joint', 'freejoint'):
          a = defaults.resolve(je, 'joint', je.attrib.get('class', cc))
          if je.tag == 'freejoint':
            a = dict(je.attrib)
          b.joints.append(_parse_joint(a, je.tag, deg))
      for ge in be.findall('geom'):
        g = _parse_geom(defaults.resolve(ge, 'geom', ge.attrib.get('class', cc)), deg=deg)
        b.geoms.append(g)
      if be.find('inertial') is not None:
        raise NotImplementedError('explicit <inertial> is not supported')
      idx = len(bodies)
      bodies.append(b)
      for g in b.geoms:
        g.body = idx
      _body_inertial(b)
      walk(be, idx, cc)

  walk(world, -1, None)
  if not bodies:
    raise ValueError('model has no bodies')
  total = float(comp.get('settotalmass', -1))
  if total > 0:   # compiler settotalmass: rescale every body's mass and inertia
    scale = total / sum(b.mass for b in bodies)
    for b in bodies:
      b.mass *= scale
      b.inertia = b.inertia * scale
  for b in bodies:
    if not b.joints:
      raise NotImplementedError('static non-world bodies are not supported')
    types = [j.type for j in b.joints]
    if 0 in types and len(types) > 1:
      raise RuntimeError('invalid joint stack: cannot stack free joints')
    if len(types) > 3:
      raise NotImplementedError('at most 3 stacked joints per body')
    if any((j.pos != b.joints[0].pos).any() for j in b.joints):
      raise RuntimeError('invalid joint stack: only one joint position allowed')
    for j in b.joints:
      if j.ref != 0:
        raise NotImplementedError(
            'The `ref` attribute on joint types is not supported.')
      if j.type == 0 and j.stiffness > 0:
        raise RuntimeError('brax does not support stiffness for free joints')

  # ---- links ----------------------------------------------------------
  nl = len(bodies)
  link_types = ''.join(
      'f' if b.joints[0].type == 0 else str(len(b.joints)) for b in bodies)
  link_parents = tuple(b.parent for b in bodies)
  for i, t in enumerate(link_types):
    if t == 'f' and link_parents[i] != -1:
      raise NotImplementedError('free joints must be on root bodies')
  nq = sum(7 if t == 'f' else int(t) for t in link_types)
  nv = sum(6 if t == 'f' else int(t) for t in link_types)

  dof_inv, body_inv, m0 = _invweight0(bodies, nv)

  tpos = np.array([b.pos for b in bodies])
  trot = np.array([b.quat for b in bodies])
  qpos0 = []
  for i, (b, t) in enumerate(zip(bodies, link_types)):
    if t == 'f':
      qpos0.extend(list(b.pos) + list(b.quat))
      tpos[i] = 0.0
      trot[i] = np.array([1.0, 0, 0, 0])
    else:
      qpos0.extend([0.0] * int(t))
  qpos0 = np.array(qpos0)

  link = base.Link(
      transform=base.Transform(pos=tpos, rot=trot),
      joint=base.Transform(
          pos=np.array([b.joints[0].pos for b in bodies]),
          rot=np.tile(np.array([1.0, 0, 0, 0]), (nl, 1))),
      inertia=base.Inertia(
          transform=base.Transform(
              pos=np.array([b.ipos for b in bodies]),
              rot=np.array([b.iquat for b in bodies])),
          i=np.array([np.diag(b.inertia) for b in bodies]),
          mass=np.array([b.mass for b in bodies])),
      invweight=body_inv[:, 0].copy(),
  )

  # ---- dofs -----------------------------------------------------------
  m_ang, m_vel, lo, hi, stiff, arm, damp, sp = [], [], [], [], [], [], [], []
  any_limited = False
  for b in bodies:
    for j in b.joints:
      params = np.concatenate([j.solref, j.solimp])
      if j.type == 0:
        m_ang.append(np.eye(6, 3, -3)); m_vel.append(np.eye(6, 3))
        lo.extend([-np.inf] * 6); hi.extend([np.inf] * 6)
        stiff.extend([0.0] * 6)
        arm.extend([j.armature] * 6); damp.extend([j.damping] * 6)
        sp.extend([params] * 6)
      else:
        if j.type == 2:
          m_ang.append(np.zeros((1, 3))); m_vel.append(j.axis.reshape(1, 3))
        else:
          m_ang.append(j.axis.reshape(1, 3)); m_vel.append(np.zeros((1, 3)))
        if j.limited:
          any_limited = True
          lo.append(j.range[0]); hi.append(j.range[1])
        else:
          lo.append(-np.inf); hi.append(np.inf)
        stiff.append(j.stiffness); arm.append(j.armature)
        damp.append(j.damping); sp.append(params)
  dof = base.DoF(
      motion=base.Motion(ang=np.concatenate(m_ang), vel=np.concatenate(m_vel)),
      armature=np.array(arm), stiffness=np.array(stiff), damping=np.array(damp),
      limit=(np.array(lo), np.array(hi)) if any_limited else None,
      invweight=dof_inv, solver_params=np.array(sp))

  # ---- actuators ------------------------------------------------------
  jnt_adr: Dict[str, Tuple[int, int]] = {}
  qa, da = 0, 0
  for b in bodies:
    for j in b.joints:
      jnt_adr[j.name] = (qa, da)
      qa += 7 if j.type == 0 else 1
      da += 6 if j.type == 0 else 1
  q_id, qd_id, gain, gear, cr, fr, bq, bqd = [], [], [], [], [], [], [], []
  act = root.find('actuator')
  for ae in (list(act) if act is not None else []):
    a = defaults.resolve(ae, ae.tag, ae.attrib.get('class'))
    if ae.tag in ('position', 'velocity', 'general'):
      a = {**defaults.resolve(ae, 'general', ae.attrib.get('class')), **a}
    if 'joint' not in a:
      raise NotImplementedError(
          'Only joint transmission types are supported for actuators.')
    qi, di = jnt_adr[a['joint']]
    q_id.append(qi); qd_id.append(di)
    gear.append(_vec(a.get('gear'), [1.0])[0])
    if ae.tag == 'motor':
      gain.append(1.0); bq.append(0.0); bqd.append(0.0)
    elif ae.tag == 'position':
      kp = float(a.get('kp', 1.0)); kv = float(a.get('kv', 0.0))
      gain.append(kp); bq.append(-kp); bqd.append(-kv)
    elif ae.tag == 'velocity':
      kv = float(a.get('kv', 1.0))
      gain.append(kv); bq.append(0.0); bqd.append(-kv)
    elif ae.tag == 'general':
      if a.get('gaintype', 'fixed') != 'fixed':
        raise NotImplementedError('Only actuator_gaintype in [0] is supported.')
      gp = _vec(a.get('gainprm'), [1.0])
      bp = _vec(a.get('biasprm'), [0.0, 0.0, 0.0])
      bp = np.concatenate([bp, np.zeros(3)])[:3]
      bt = a.get('biastype', 'none')
      if bt not in ('none', 'affine'):
        raise NotImplementedError(
            'Only actuator_biastype in [0, 1] are supported.')
      on = 0.0 if bt == 'none' else 1.0
      gain.append(gp[0]); bq.append(bp[1] * on); bqd.append(bp[2] * on)
    else:
      raise NotImplementedError(f'actuator <{ae.tag}> is not supported')
    ctrl = _vec(a.get('ctrlrange'), [0, 0])
    cl = a.get('ctrllimited', 'auto')
    cl = ('ctrlrange' in a) if cl == 'auto' else cl == 'true'
    cr.append(ctrl if cl else np.array([-np.inf, np.inf]))
    frc = _vec(a.get('forcerange'), [0, 0])
    fl = a.get('forcelimited', 'auto')
    fl = ('forcerange' in a) if fl == 'auto' else fl == 'true'
    fr.append(frc if fl else np.array([-np.inf, np.inf]))
  nu = len(q_id)
  actuator = base.Actuator(
      q_id=np.array(q_id, dtype=np.int32), qd_id=np.array(qd_id, dtype=np.int32),
      ctrl_range=np.array(cr).reshape(nu, 2), force_range=np.array(fr).reshape(nu, 2),
      gain=np.array(gain), gear=np.array(gear),
      bias_q=np.array(bq), bias_qd=np.array(bqd))

  # ---- geoms (world geoms first, then bodies in order, as MuJoCo does) --
  geoms = world_geoms + [g for b in bodies for g in b.geoms]
  if geoms:
    if any(g.solmix != geoms[0].solmix for g in geoms):
      raise NotImplementedError('geom_solmix parameter not supported.')
    if any(g.priority != geoms[0].priority for g in geoms):
      raise NotImplementedError('geom_priority parameter not supported.')

  init_q = custom['init_qpos'] if 'init_qpos' in custom else qpos0
  if init_q.shape[0] != nq:
    raise ValueError(
        f'init_qpos had length {init_q.shape[0]} but expected length {nq}.')

  sys = base.System(
      gravity=gravity, viscosity=viscosity, density=density,
      link=link, dof=dof, actuator=actuator, init_q=init_q,
      opt=base.Option(timestep=timestep, iterations=iterations),
      enable_fluid=bool(viscosity > 0 or density > 0),
      link_names=[b.name for b in bodies],
      link_types=link_types, link_parents=link_parents,
      matrix_inv_iterations=int(custom.get(
          'matrix_inv_iterations', [_CUSTOM_DEFAULTS['matrix_inv_iterations']])[0]),
      solver_iterations=iterations,
      solver_maxls=int(custom.get(
          'solver_maxls', [_CUSTOM_DEFAULTS['solver_maxls']])[0]),
      nq=nq, nv=nv, nu=nu,
      geom_bodyid=np.array([g.body + 1 for g in geoms], dtype=np.int32),
      geom_type=np.array([g.type for g in geoms], dtype=np.int32),
      geom_pos=np.array([g.pos for g in geoms]).reshape(-1, 3),
      geom_quat=np.array([g.quat for g in geoms]).reshape(-1, 4),
      geom_size=np.array([g.size for g in geoms]).reshape(-1, 3),
      geom_friction=np.array([g.friction for g in geoms]).reshape(-1, 3),
      geom_solref=np.array([g.solref for g in geoms]).reshape(-1, 2),
      geom_solimp=np.array([g.solimp for g in geoms]).reshape(-1, 5),
      geom_contype=np.array([g.contype for g in geoms], dtype=np.int32),
      geom_conaffinity=np.array([g.conaffinity for g in geoms], dtype=np.int32),
      geom_names=[g.name for g in geoms],
      qpos0=qpos0, mass_mx0=m0,
  )
  return sys.cast()


def load(path) -> base.System:
  """Loads a System from an MJCF file path (reference `io/mjcf.py:525-528`)."""
  with open(path, 'r') as f:
    return loads(f.read())
