"""Core types: the `System` / `State` containers the generalized step consumes.

These mirror the field set of the reference pytrees so that callers written
against `brax.base` find the same attribute paths:

  Transform / Motion / Force / Inertia   reference `brax/base.py:170-302`
  Link / DoF / Actuator                  reference `brax/base.py:305-393`
  State                                  reference `brax/base.py:396-412`
  System                                 reference `brax/base.py:415-540`

Leaves are NumPy arrays (model constants) or torch tensors (batched device
state).  There is no tracing compiler here, so "pytree" support is a small
dataclass walker: `tree_map`, `.replace`, `.take`, `.tree_replace`.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

Q_WIDTHS = {'f': 7, '1': 1, '2': 2, '3': 3}
QD_WIDTHS = {'f': 6, '1': 1, '2': 2, '3': 3}


def _is_node(x) -> bool:
  return dataclasses.is_dataclass(x) and not isinstance(x, type)


def tree_map(fn: Callable, tree, *rest):
  """Maps `fn` over array leaves of (nested) dataclasses / tuples / lists."""
  if tree is None:
    return None
  if _is_node(tree):
    kw = {}
    for f in dataclasses.fields(tree):
      v = getattr(tree, f.name)
      if f.metadata.get('static', False):
        kw[f.name] = v
      else:
        kw[f.name] = tree_map(fn, v, *[getattr(r, f.name) for r in rest])
    return type(tree)(**kw)
  if isinstance(tree, (tuple, list)):
    return type(tree)(
        tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
  return fn(tree, *rest)


def tree_leaves(tree) -> List[Any]:
  out: List[Any] = []
  tree_map(lambda x: out.append(x) or x, tree)
  return out


def unbatch(tree_v, in_axes) -> list:
  """The n unbatched trees a batched tree stands for: leaf i of result e is `leaf[e]` where `in_axes` holds 0 and the
  shared leaf where it holds None.  This is what `jax.vmap(f, in_axes=[in_axes, ...])(tree_v, ...)` hands to `f` for
  env e (reference envs/wrappers/training.py:223-260: a System with domain-randomised leaves).  `in_axes` has the
  structure of `tree_v`; a None at a node stands for None at every leaf below it."""
  sizes = set()

  def scan(t, ax):
    if t is None or ax is None:
      return
    if _is_node(t):
      for f in dataclasses.fields(t):
        if not f.metadata.get('static', False):
          scan(getattr(t, f.name), getattr(ax, f.name) if _is_node(ax) else ax)
    elif isinstance(t, (tuple, list)):
      for i, v in enumerate(t):
        scan(v, ax[i] if isinstance(ax, (tuple, list)) else ax)
    else:
      if ax != 0:
        raise NotImplementedError(f'in_axes entries must be 0 or None, got {ax!r}')
      sizes.add(int(np.shape(t)[0]))
  scan(tree_v, in_axes)
  if len(sizes) != 1:
    raise ValueError(f'batched leaves must share one leading size, got {sorted(sizes)}')
  n = sizes.pop()

  def pick(t, ax, e):
    if t is None or ax is None:
      return t
    if _is_node(t):
      kw = {}
      for f in dataclasses.fields(t):
        v = getattr(t, f.name)
        kw[f.name] = v if f.metadata.get('static', False) else pick(v, getattr(ax, f.name) if _is_node(ax) else ax, e)
      return type(t)(**kw)
    if isinstance(t, (tuple, list)):
      return type(t)(pick(v, ax[i] if isinstance(ax, (tuple, list)) else ax, e) for i, v in enumerate(t))
    return t[e]
  return [pick(tree_v, in_axes, e) for e in range(n)]


def static(default=dataclasses.MISSING, **kw):
  if default is dataclasses.MISSING:
    return dataclasses.field(metadata={'static': True}, **kw)
  return dataclasses.field(default=default, metadata={'static': True}, **kw)


class Base:
  """Array-like helpers shared by all containers (reference `base.py:38-158`)."""

  def replace(self, **kw):
    return dataclasses.replace(self, **kw)

  def __add__(self, o):
    return tree_map(lambda x, y: x + y, self, o)

  def __sub__(self, o):
    return tree_map(lambda x, y: x - y, self, o)

  def __mul__(self, o):
    return tree_map(lambda x: x * o, self)

  def __neg__(self):
    return tree_map(lambda x: -x, self)

  def __truediv__(self, o):
    return tree_map(lambda x: x / o, self)

  def take(self, i, axis=0):
    def f(x):
      if isinstance(x, np.ndarray):
        return np.take(x, i, axis=axis, mode='wrap')
      import torch
      idx = torch.as_tensor(i, device=x.device)
      return torch.index_select(x, axis, idx.reshape(-1) % x.shape[axis]).reshape(
          x.shape[:axis] + tuple(idx.shape) + x.shape[axis + 1:])
    return tree_map(f, self)

  def tree_replace(self, params: Dict[str, Any]):
    """`sys.tree_replace({'opt.timestep': dt})` (reference `base.py:118-137`)."""
    new = self
    for k, v in params.items():
      new = _tree_replace(new, k.split('.'), v)
    return new


def _tree_replace(node, attr: Sequence[str], val):
  if not attr:
    return node
  if len(attr) == 1:
    return node.replace(**{attr[0]: val})
  return node.replace(
      **{attr[0]: _tree_replace(getattr(node, attr[0]), attr[1:], val)})


@dataclasses.dataclass(frozen=True)
class Transform(Base):
  """pos (…,3) and unit-quaternion rot (…,4), (w,x,y,z)."""
  pos: Any
  rot: Any


@dataclasses.dataclass(frozen=True)
class Motion(Base):
  """Spatial motion: ang (…,3), vel (…,3)."""
  ang: Any
  vel: Any


@dataclasses.dataclass(frozen=True)
class Force(Base):
  ang: Any
  vel: Any


@dataclasses.dataclass(frozen=True)
class Inertia(Base):
  """transform: inertial frame; i: (…,3,3); mass: (…,)."""
  transform: Transform
  i: Any
  mass: Any


@dataclasses.dataclass(frozen=True)
class Link(Base):
  transform: Transform
  joint: Transform
  inertia: Inertia
  invweight: Any


@dataclasses.dataclass(frozen=True)
class DoF(Base):
  motion: Motion
  armature: Any
  stiffness: Any
  damping: Any
  limit: Optional[Tuple[Any, Any]]
  invweight: Any
  solver_params: Any


@dataclasses.dataclass(frozen=True)
class Actuator(Base):
  q_id: Any
  qd_id: Any
  ctrl_range: Any
  force_range: Any
  gain: Any
  gear: Any
  bias_q: Any
  bias_qd: Any


@dataclasses.dataclass(frozen=True)
class Option(Base):
  """The two `mjx.Model.opt` members the hot path reads."""
  timestep: Any
  iterations: int = static(100)


@dataclasses.dataclass(frozen=True)
class State(Base):
  """reference `brax/base.py:396-412`."""
  q: Any
  qd: Any
  x: Transform
  xd: Motion
  contact: Any


@dataclasses.dataclass(frozen=True)
class ContactPairs:
  """Static collision-pair table derived from the geoms (world link = -1).

  Restates what `mjx.make_data(sys).ncon` / `mjx.collision` enumerate for the
  supported pair types: contype/conaffinity filter, same-body and parent-child
  exclusion.  plane-sphere gives one contact per pair; plane-capsule gives two
  (the capsule's end spheres, +axis first), emitted after the plane-sphere group
  (mjx groups contacts by geom-type pair).  For capsules `frame` is only the
  fallback: the tangent follows the capsule axis at run time.  capsule-capsule
  (kind 2, one contact between the closest points of the two segments) comes last;
  both geoms may sit on moving links and geom1's shape is in the `a_*` fields.
  """
  geom1: np.ndarray        # (ncon,) plane geom
  geom2: np.ndarray        # (ncon,) sphere / capsule geom
  kind: np.ndarray         # (ncon,) 0 = plane-sphere, 1 = plane-capsule end point
  geom_quat: np.ndarray    # (ncon,4) capsule orientation in link_b frame
  half_len: np.ndarray     # (ncon,) signed half length: end point = centre + axis * half_len
  link_a: np.ndarray       # (ncon,) link of geom1 (-1 = world)
  link_b: np.ndarray       # (ncon,) link of geom2
  plane_pos: np.ndarray    # (ncon,3) world
  plane_normal: np.ndarray  # (ncon,3) world (plane z axis)
  frame: np.ndarray        # (ncon,3,3) rows n, t1, t2 (mjx math.make_frame)
  sphere_pos: np.ndarray   # (ncon,3) in link_b frame
  radius: np.ndarray       # (ncon,)
  friction: np.ndarray     # (ncon,) sliding friction mu = max(geom1, geom2)
  solref: np.ndarray       # (ncon,2)
  solimp: np.ndarray       # (ncon,5)
  a_pos: np.ndarray = None     # (ncon,3) kind 2: geom1 centre in link_a frame
  a_quat: np.ndarray = None    # (ncon,4) kind 2: geom1 orientation in link_a frame
  a_half: np.ndarray = None    # (ncon,)  kind 2: geom1 half length
  a_radius: np.ndarray = None  # (ncon,)  kind 2: geom1 radius


@dataclasses.dataclass(frozen=True)
class System(Base):
  """reference `brax/base.py:415-540` (hot-path subset + geoms)."""
  gravity: Any
  viscosity: Any
  density: Any
  link: Link
  dof: DoF
  actuator: Actuator
  init_q: Any
  opt: Option
  geom_pos: Any
  geom_quat: Any
  geom_size: Any
  geom_friction: Any
  geom_solref: Any
  geom_solimp: Any
  qpos0: Any
  mass_mx0: Any
  enable_fluid: bool = static(False)
  link_names: List[str] = static(default_factory=list)
  link_types: str = static('')
  link_parents: Tuple[int, ...] = static(())
  matrix_inv_iterations: int = static(10)
  solver_iterations: int = static(100)
  solver_maxls: int = static(20)
  nq: int = static(0)
  nv: int = static(0)
  nu: int = static(0)
  geom_bodyid: Any = static(None)
  geom_type: Any = static(None)
  geom_contype: Any = static(None)
  geom_conaffinity: Any = static(None)
  geom_names: List[str] = static(default_factory=list)

  # -- reference methods ----------------------------------------------------
  def num_links(self) -> int:
    return len(self.link_types)

  def dof_link(self, depth: bool = False) -> np.ndarray:
    idx: List[int] = []
    for i, t in enumerate(self.link_types):
      idx.extend([i] * QD_WIDTHS[t])
    if depth:
      count: Dict[int, int] = {}
      per_link = []
      for i in range(self.num_links()):
        d = self.link_depth(i)
        per_link.append(count.get(d, 0))
        count[d] = count.get(d, 0) + 1
      idx = [per_link[i] for i in idx]
    return np.array(idx, dtype=np.int32)

  def link_depth(self, i: int) -> int:
    d = 0
    while self.link_parents[i] >= 0:
      i = self.link_parents[i]
      d += 1
    return d

  def dof_ranges(self) -> List[List[int]]:
    beg, out = 0, []
    for t in self.link_types:
      out.append(list(range(beg, beg + QD_WIDTHS[t])))
      beg += QD_WIDTHS[t]
    return out

  def q_idx(self, link_type: str) -> np.ndarray:
    i, out = 0, []
    for t in self.link_types:
      if t in link_type:
        out.extend(range(i, i + Q_WIDTHS[t]))
      i += Q_WIDTHS[t]
    return np.array(out, dtype=np.int32)

  def qd_idx(self, link_type: str) -> np.ndarray:
    i, out = 0, []
    for t in self.link_types:
      if t in link_type:
        out.extend(range(i, i + QD_WIDTHS[t]))
      i += QD_WIDTHS[t]
    return np.array(out, dtype=np.int32)

  def q_size(self) -> int:
    return self.nq

  def qd_size(self) -> int:
    return self.nv

  def act_size(self) -> int:
    return self.nu

  # -- helpers ---------------------------------------------------------------
  def cast(self) -> 'System':
    """float64 -> float32 for every float leaf (reference `io/mjcf.py:478`)."""
    def f(x):
      if isinstance(x, np.ndarray) and x.dtype == np.float64:
        return x.astype(np.float32)
      if isinstance(x, float):
        return np.float32(x)
      return x
    return tree_map(f, self)

  def contact_pairs(self) -> ContactPairs:
    """Enumerates colliding geom pairs (see ContactPairs)."""
    ng = 0 if self.geom_type is None else len(self.geom_type)
    rows, caps, capcap = [], [], []
    for g1 in range(ng):
      for g2 in range(g1 + 1, ng):
        b1, b2 = int(self.geom_bodyid[g1]), int(self.geom_bodyid[g2])
        ok = (self.geom_contype[g1] & self.geom_conaffinity[g2]) or (
            self.geom_contype[g2] & self.geom_conaffinity[g1])
        if not ok:
          continue
        if b1 == b2:
          continue
        l1, l2 = b1 - 1, b2 - 1
        # parent-child filter (world-child pairs are kept, as in MuJoCo)
        if l1 >= 0 and l2 >= 0 and (
            self.link_parents[l1] == l2 or self.link_parents[l2] == l1):
          continue
        t1, t2 = int(self.geom_type[g1]), int(self.geom_type[g2])
        if (t1, t2) == (3, 3):
          capcap.append((g1, g2, l1, l2, 2, float(self.geom_size[g2][1])))
          continue
        if (t1, t2) not in ((0, 2), (0, 3)):
          raise NotImplementedError(
              f'collision pair type ({t1},{t2}) not supported: only plane-sphere, plane-capsule '
              'and capsule-capsule (SURVEY.md section 8 a-11 / f-3)')
        if l1 != -1:
          raise NotImplementedError('planes must be attached to the world')
        if t2 == 2:
          rows.append((g1, g2, l1, l2, 0, 0.0))
        else:
          half = float(self.geom_size[g2][1])
          caps.append((g1, g2, l1, l2, 1, half))
          caps.append((g1, g2, l1, l2, 1, -half))
    rows = rows + caps + capcap
    n = len(rows)
    f64 = np.float64

    def quat_to_mat(q):
      w, x, y, z = (q / np.linalg.norm(q)).astype(f64)
      return np.array([
          [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
          [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
          [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    def make_frame(a):
      a = a / np.linalg.norm(a)
      b = np.array([0.0, 1.0, 0.0]) if -0.5 < a[1] < 0.5 else np.array([0.0, 0.0, 1.0])
      b = b - a * a.dot(b)
      b = b / np.linalg.norm(b)
      return np.stack([a, b, np.cross(a, b)])

    cp = dict(
        kind=np.zeros(n, np.int32), geom_quat=np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)),
        half_len=np.zeros(n, np.float32),
        geom1=np.zeros(n, np.int32), geom2=np.zeros(n, np.int32),
        link_a=np.zeros(n, np.int32), link_b=np.zeros(n, np.int32),
        plane_pos=np.zeros((n, 3), np.float32),
        plane_normal=np.zeros((n, 3), np.float32),
        frame=np.zeros((n, 3, 3), np.float32),
        sphere_pos=np.zeros((n, 3), np.float32), radius=np.zeros(n, np.float32),
        friction=np.zeros(n, np.float32), solref=np.zeros((n, 2), np.float32),
        solimp=np.zeros((n, 5), np.float32),
        a_pos=np.zeros((n, 3), np.float32), a_quat=np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)),
        a_half=np.zeros(n, np.float32), a_radius=np.zeros(n, np.float32))
    for k, (g1, g2, l1, l2, kind, half) in enumerate(rows):
      cp['kind'][k], cp['half_len'][k] = kind, half
      cp['geom_quat'][k] = self.geom_quat[g2]
      nrm = quat_to_mat(np.asarray(self.geom_quat[g1], f64))[:, 2]
      cp['geom1'][k], cp['geom2'][k] = g1, g2
      cp['link_a'][k], cp['link_b'][k] = l1, l2
      if kind == 2:
        cp['a_pos'][k], cp['a_quat'][k] = self.geom_pos[g1], self.geom_quat[g1]
        cp['a_half'][k], cp['a_radius'][k] = self.geom_size[g1][1], self.geom_size[g1][0]
      else:
        cp['plane_pos'][k] = self.geom_pos[g1]
        cp['plane_normal'][k] = nrm
        cp['frame'][k] = make_frame(nrm)
      cp['sphere_pos'][k] = self.geom_pos[g2]
      cp['radius'][k] = self.geom_size[g2][0]
      cp['friction'][k] = max(self.geom_friction[g1][0], self.geom_friction[g2][0])
      # equal solmix / priority (validated at load): plain average
      cp['solref'][k] = 0.5 * (np.asarray(self.geom_solref[g1], f64)
                               + np.asarray(self.geom_solref[g2], f64))
      cp['solimp'][k] = 0.5 * (np.asarray(self.geom_solimp[g1], f64)
                               + np.asarray(self.geom_solimp[g2], f64))
    return ContactPairs(**cp)
