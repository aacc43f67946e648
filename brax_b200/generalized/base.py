"""generalized.State -- same leaves as reference `brax/generalized/base.py:25-92`.

Every leaf is a torch CUDA tensor with a leading env axis (the reference adds
that axis with `jax.vmap`; here the kernel is natively batched)."""
from __future__ import annotations

import dataclasses
from typing import Any, Dict

from brax_b200 import base
from brax_b200.base import Inertia, Motion, Transform


@dataclasses.dataclass(frozen=True)
class State(base.State):
  """Dynamic state that changes after every step (reference field set)."""
  root_com: Any
  cinr: Inertia
  cd: Motion
  cdof: Motion
  cdofd: Motion
  mass_mx: Any
  mass_mx_inv: Any
  con_jac: Any
  con_diag: Any
  con_aref: Any
  qf_smooth: Any
  qf_constraint: Any
  qdd: Any

  # ---- flat <-> nested ----------------------------------------------------
  @classmethod
  def from_flat(cls, b: Dict[str, Any], contact=None) -> 'State':
    # a lean state (BXG_STEP_LEAN, include/bxg.h) carries q, qd, x, xd and mass_mx_inv only: the other leaves are None
    g = b.get
    return cls(
        q=b['q'], qd=b['qd'],
        x=Transform(pos=b['x_pos'], rot=b['x_rot']),
        xd=Motion(ang=b['xd_ang'], vel=b['xd_vel']),
        contact=contact,
        root_com=g('root_com'),
        cinr=Inertia(transform=Transform(pos=g('cinr_pos'), rot=g('cinr_rot')),
                     i=g('cinr_i'), mass=g('cinr_mass')),
        cd=Motion(ang=g('cd_ang'), vel=g('cd_vel')),
        cdof=Motion(ang=g('cdof_ang'), vel=g('cdof_vel')),
        cdofd=Motion(ang=g('cdofd_ang'), vel=g('cdofd_vel')),
        mass_mx=g('mass_mx'), mass_mx_inv=b['mass_mx_inv'],
        con_jac=g('con_jac'), con_diag=g('con_diag'), con_aref=g('con_aref'),
        qf_smooth=g('qf_smooth'), qf_constraint=g('qf_constraint'), qdd=g('qdd'))

  def to_flat(self) -> Dict[str, Any]:
    return {
        'q': self.q, 'qd': self.qd, 'x_pos': self.x.pos, 'x_rot': self.x.rot,
        'xd_ang': self.xd.ang, 'xd_vel': self.xd.vel, 'root_com': self.root_com,
        'cinr_pos': self.cinr.transform.pos, 'cinr_rot': self.cinr.transform.rot,
        'cinr_i': self.cinr.i, 'cinr_mass': self.cinr.mass,
        'cd_ang': self.cd.ang, 'cd_vel': self.cd.vel,
        'cdof_ang': self.cdof.ang, 'cdof_vel': self.cdof.vel,
        'cdofd_ang': self.cdofd.ang, 'cdofd_vel': self.cdofd.vel,
        'mass_mx': self.mass_mx, 'mass_mx_inv': self.mass_mx_inv,
        'con_jac': self.con_jac, 'con_diag': self.con_diag, 'con_aref': self.con_aref,
        'qf_smooth': self.qf_smooth, 'qf_constraint': self.qf_constraint, 'qdd': self.qdd}
