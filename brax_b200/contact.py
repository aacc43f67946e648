"""`contact.get(sys, x)` for `debug=True` (reference `brax/contact.py:28-67`).

The reference returns `mjx.collision`'s Contact (plus `link_idx`, `elasticity`) when `pipeline.init / step` run with
`debug=True` (`generalized/pipeline.py:58-60,91-92`).  The step kernel evaluates the same colliders on chip for the
constraint Jacobian and only reports the penetration distances (`BxgDiag.con_dist`); this module restates the three
supported colliders -- plane-sphere, plane-capsule (two end points), capsule-capsule -- as batched torch ops on the
state's link transforms, for inspection and visualisation.  It is debug tooling: nothing on the step path calls it.
Geometry follows mjx `collision_primitive.py` (`plane_sphere`, `plane_capsule`, `capsule_capsule`) and
`math.closest_segment_to_segment_points`, as `oracle/bxg_oracle.c` states them.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Optional

import numpy as np
import torch

from brax_b200 import base


@dataclasses.dataclass(frozen=True)
class Contact(base.Base):
  """reference `brax/base.py:359-368`: mjx.Contact + link_idx + elasticity.  Leading axis: envs.

  dist (n, ncon), pos (n, ncon, 3), frame (n, ncon, 3, 3) rows normal / tangent / bitangent, includemargin (ncon,),
  friction (ncon, 5), solref (ncon, 2), solimp (ncon, 5), geom1 / geom2 (ncon,), link_idx ((ncon,), (ncon,)),
  elasticity (ncon,).  `solver_stats` (n, 4) is this backend's addition: projected-gradient iterations, line-search
  trials, Newton-Schulz accepts and cold starts of the step that produced the state (None after `init`)."""
  dist: Any
  pos: Any
  frame: Any
  includemargin: Any
  friction: Any
  solref: Any
  solimp: Any
  geom1: Any
  geom2: Any
  link_idx: Any
  elasticity: Any
  solver_stats: Any = None

  def __getitem__(self, key):   # the names of the kernel's diagnostics (BxgDiag)
    return {'con_dist': self.dist, 'stats': self.solver_stats}[key]


def _rotate(v, q):
  s, u = q[..., :1], q[..., 1:]
  return 2 * ((u * v).sum(-1, keepdim=True) * u) + (s * s - (u * u).sum(-1, keepdim=True)) * v + 2 * s * torch.cross(u, v, dim=-1)


def _quat_mul(a, b):
  aw, ax, ay, az = a.unbind(-1)
  bw, bx, by, bz = b.unbind(-1)
  return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                      aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def _z_axis(q):
  """third column of quat_to_3x3(q)"""
  w, x, y, z = q.unbind(-1)
  d = (q * q).sum(-1)
  s = 2.0 / d
  return torch.stack([(x * z + w * y) * s, (y * z - w * x) * s, 1.0 - (x * x + y * y) * s], -1)


def _normalize(v, eps=0.0):
  n = torch.linalg.vector_norm(v, dim=-1, keepdim=True)
  return v / torch.where(n > eps, n, torch.ones_like(n)), n[..., 0]


def _make_frame(a):
  """mjx math.make_frame: rows a, b, a x b with b = y (z when a is near +-y) made orthogonal to a."""
  a, _ = _normalize(a)
  use_y = ((a[..., 1] > -0.5) & (a[..., 1] < 0.5))[..., None]
  y = torch.tensor([0.0, 1.0, 0.0], dtype=a.dtype, device=a.device).expand_as(a)
  z = torch.tensor([0.0, 0.0, 1.0], dtype=a.dtype, device=a.device).expand_as(a)
  b = torch.where(use_y, y, z)
  b = b - a * (a * b).sum(-1, keepdim=True)
  b, _ = _normalize(b)
  return torch.stack([a, b, torch.cross(a, b, dim=-1)], -2)


def _closest_segment_point(a, b, pt):
  ab = b - a
  t = ((pt - a) * ab).sum(-1, keepdim=True) / ((ab * ab).sum(-1, keepdim=True) + 1e-6)
  return a + t.clamp(0.0, 1.0) * ab


def _closest_segment_to_segment_points(a0, a1, b0, b1):
  dir_a, len_a = _normalize(a1 - a0)
  dir_b, len_b = _normalize(b1 - b0)
  half_a, half_b = (0.5 * len_a)[..., None], (0.5 * len_b)[..., None]
  a_mid, b_mid = a0 + dir_a * half_a, b0 + dir_b * half_b
  trans = a_mid - b_mid
  dab, dat, dbt = (dir_a * dir_b).sum(-1, keepdim=True), (dir_a * trans).sum(-1, keepdim=True), (dir_b * trans).sum(-1, keepdim=True)
  t_a0 = (-dat + dab * dbt) / ((1 - dab * dab) + 1e-6)
  t_b0 = dbt + t_a0 * dab
  t_a = torch.maximum(-half_a, torch.minimum(t_a0, half_a))
  t_b = torch.maximum(-half_b, torch.minimum(t_b0, half_b))
  best_a, best_b = a_mid + dir_a * t_a, b_mid + dir_b * t_b
  new_a, new_b = _closest_segment_point(a0, a1, best_b), _closest_segment_point(b0, b1, best_a)
  d1, d2 = ((new_a - best_b) ** 2).sum(-1, keepdim=True), ((best_a - new_b) ** 2).sum(-1, keepdim=True)
  return torch.where(d1 < d2, new_a, best_a), torch.where(d1 < d2, best_b, new_b)


def get(sys: base.System, x: base.Transform, kernel_dist=None, solver_stats=None) -> Optional[Contact]:
  """Calculates contacts from the link transforms `x` (pos (n, L, 3), rot (n, L, 4)); None if the System has none.

  kernel_dist: the distances the step kernel reported for this state (`BxgDiag.con_dist`); they replace the ones
  computed here (same quantity, float32 on both sides)."""
  cp = sys.contact_pairs()
  ncon = len(cp.geom1)
  if ncon == 0:
    if solver_stats is None:
      return None
    e = torch.zeros((x.pos.shape[0] if x.pos.dim() == 3 else 1, 0), dtype=x.pos.dtype, device=x.pos.device)   # (the step's counters still travel)
    return Contact(e, e[..., None].expand(-1, 0, 3), e[..., None, None].expand(-1, 0, 3, 3), e[0], e[0, :, None].expand(0, 5), e[0, :, None].expand(0, 2),
                   e[0, :, None].expand(0, 5), e[0].long(), e[0].long(), (e[0].long(), e[0].long()), e[0], solver_stats)
  pos, rot = x.pos, x.rot
  if pos.dim() == 2:
    pos, rot = pos[None], rot[None]
  n, dev, dt = pos.shape[0], pos.device, pos.dtype
  T = lambda a, d=None: torch.as_tensor(np.asarray(a), device=dev).to(d or dt)   # noqa: E731
  kind = np.asarray(cp.kind)
  lb = torch.as_tensor(np.asarray(cp.link_b, np.int64), device=dev)
  bp, br = pos[:, lb], rot[:, lb]                                  # (n, ncon, 3 | 4)
  centre = bp + _rotate(T(cp.sphere_pos)[None].expand_as(bp), br)  # sphere / capsule centre
  radius = T(cp.radius)[None]
  nrm = T(cp.plane_normal)[None].expand(n, ncon, 3)
  frame = T(cp.frame)[None].expand(n, ncon, 3, 3).clone()
  axis_b = _z_axis(_quat_mul(br, T(cp.geom_quat)[None].expand_as(br)))
  half_b = T(cp.half_len)[None, :, None]
  # plane-sphere and plane-capsule end points: a sphere against the plane
  end = centre + axis_b * half_b * T(kind == 1)[None, :, None]
  d_plane = ((end - T(cp.plane_pos)[None]) * nrm).sum(-1) - radius
  pos_plane = end - nrm * (radius + 0.5 * d_plane)[..., None]
  # plane-capsule: the tangent follows the capsule axis (fallback: make_frame's axis when the capsule stands on end)
  b = axis_b - nrm * (nrm * axis_b).sum(-1, keepdim=True)
  b, bn = _normalize(b)
  fb = _make_frame(nrm)[..., 1, :]
  b = torch.where((bn < 0.5)[..., None], fb, b)
  cap = T(kind == 1, torch.bool)[None, :, None]
  frame[..., 1, :] = torch.where(cap, b, frame[..., 1, :])
  frame[..., 2, :] = torch.where(cap, torch.cross(nrm, b, dim=-1), frame[..., 2, :])
  dist, cpos = d_plane, pos_plane
  if (kind == 2).any():
    la = np.asarray(cp.link_a, np.int64)
    la_t = torch.as_tensor(np.where(la >= 0, la, 0), device=dev)
    world = T(la < 0, torch.bool)[None, :, None]
    ap = torch.where(world, torch.zeros_like(bp), pos[:, la_t])
    ar = torch.where(world, torch.tensor([1.0, 0, 0, 0], dtype=dt, device=dev).expand_as(br), rot[:, la_t])
    ca = ap + _rotate(T(cp.a_pos)[None].expand_as(ap), ar)
    axis_a = _z_axis(_quat_mul(ar, T(cp.a_quat)[None].expand_as(ar)))
    sa, sb = axis_a * T(cp.a_half)[None, :, None], axis_b * half_b
    pa, pb = _closest_segment_to_segment_points(ca - sa, ca + sa, centre - sb, centre + sb)
    nn, d = _normalize(pb - pa)
    nn = torch.where((d == 0)[..., None], torch.tensor([1.0, 0, 0], dtype=dt, device=dev).expand_as(nn), nn)
    r1 = T(cp.a_radius)[None]
    d = d - (r1 + radius)
    pos_cc = pa + nn * (r1 + 0.5 * d)[..., None]
    cc = T(kind == 2, torch.bool)[None]
    dist = torch.where(cc, d, dist)
    cpos = torch.where(cc[..., None], pos_cc, cpos)
    frame = torch.where(cc[..., None, None], _make_frame(nn), frame)
  if kernel_dist is not None:
    dist = kernel_dist
  g1, g2 = np.asarray(cp.geom1), np.asarray(cp.geom2)
  gf = np.asarray(sys.geom_friction, np.float32).reshape(-1, 3)
  f = np.maximum(gf[g1], gf[g2])
  friction = np.stack([f[:, 0], f[:, 0], f[:, 1], f[:, 2], f[:, 2]], 1)
  body = np.asarray(sys.geom_bodyid)
  return Contact(dist=dist, pos=cpos, frame=frame, includemargin=torch.zeros(ncon, dtype=dt, device=dev), friction=T(friction),
                 solref=T(cp.solref), solimp=T(cp.solimp), geom1=T(g1, torch.int64), geom2=T(g2, torch.int64),
                 link_idx=(T(body[g1] - 1, torch.int64), T(body[g2] - 1, torch.int64)),
                 elasticity=torch.zeros(ncon, dtype=dt, device=dev), solver_stats=solver_stats)
