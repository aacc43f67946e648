"""Golden vectors from the REFERENCE'S OWN SOURCE (run in the build container only).

Imports `brax.generalized.pipeline` unmodified from /root/reference and executes it on
NumPy through the stand-ins in tools/refshim/ (float64).  What is the reference's code and
what is restated (the mjx colliders, jaxopt's projected gradient, the MuJoCo model
compiler) is listed in tools/refshim/README.md.  Writes
  tests/golden/ref_<model>.npz      inputs (q0, qd0, act) and every `generalized.State` leaf after
                                    `init` and after each `step`
  tests/golden/ref_env_<env>.npz    `brax.envs.<env>` wrapped by `training.wrap` (Vmap + Episode +
                                    AutoReset): obs, reward, done, metrics, info and q, qd per env step
tests/test_reference_golden.py replays the inputs through the oracles.

  python tools/gen_reference_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brax_b200 import envs_assets   # noqa: E402  (compiled System constants: float32 values)

sys.path.insert(0, os.path.join(ROOT, 'tools', 'refshim'))   # jax / flax / jaxopt / mujoco stand-ins
sys.path.insert(0, '/root/reference')
# brax.io pulls MuJoCo's XML compiler and etils: not on the physics path
io_stub, mjcf_stub = types.ModuleType('brax.io'), types.ModuleType('brax.io.mjcf')
mjcf_stub.validate_model = lambda m: None
io_stub.mjcf = mjcf_stub
sys.modules['brax.io'], sys.modules['brax.io.mjcf'] = io_stub, mjcf_stub
# the other pipelines and the renderer are imported by brax/envs/base.py but never called here
for _name in ('brax.io.image', 'brax.mjx', 'brax.mjx.pipeline', 'brax.positional', 'brax.positional.pipeline',
              'brax.spring', 'brax.spring.pipeline'):
  sys.modules[_name] = types.ModuleType(_name)
io_stub.image = sys.modules['brax.io.image']
sys.modules['brax.mjx'].pipeline = sys.modules['brax.mjx.pipeline']
sys.modules['brax.positional'].pipeline = sys.modules['brax.positional.pipeline']
sys.modules['brax.spring'].pipeline = sys.modules['brax.spring.pipeline']

import jax.numpy as jp                                # noqa: E402  (the stand-in)
from mujoco import mjx                                # noqa: E402  (the stand-in)
from brax import base as rb                           # noqa: E402  (the reference)
from brax.generalized import pipeline as ref_pipeline  # noqa: E402  (the reference)

assert ref_pipeline.__file__.startswith('/root/reference/'), ref_pipeline.__file__

F = lambda x: jp.array(np.asarray(x, np.float64))     # noqa: E731
I = lambda x: jp.array(np.asarray(x, np.int32))       # noqa: E731


def reference_system(s):
  """brax.base.System (reference class) from our compiled constants."""
  L, nv = s.num_links(), s.nv
  z = jp.zeros(L)
  link = rb.Link(
      transform=rb.Transform(pos=F(s.link.transform.pos), rot=F(s.link.transform.rot)),
      joint=rb.Transform(pos=F(s.link.joint.pos), rot=F(s.link.joint.rot)),
      inertia=rb.Inertia(transform=rb.Transform(pos=F(s.link.inertia.transform.pos), rot=F(s.link.inertia.transform.rot)),
                         i=F(s.link.inertia.i), mass=F(s.link.inertia.mass)),
      invweight=F(s.link.invweight), constraint_stiffness=z, constraint_vel_damping=z,
      constraint_limit_stiffness=z, constraint_ang_damping=z)
  dof = rb.DoF(
      motion=rb.Motion(ang=F(s.dof.motion.ang), vel=F(s.dof.motion.vel)), armature=F(s.dof.armature),
      stiffness=F(s.dof.stiffness), damping=F(s.dof.damping),
      limit=None if s.dof.limit is None else (F(s.dof.limit[0]), F(s.dof.limit[1])),
      invweight=F(s.dof.invweight), solver_params=F(s.dof.solver_params))
  a = s.actuator
  act = rb.Actuator(q_id=I(a.q_id), qd_id=I(a.qd_id), ctrl_range=F(np.asarray(a.ctrl_range).reshape(-1, 2)),
                    force_range=F(np.asarray(a.force_range).reshape(-1, 2)), gain=F(a.gain), gear=F(a.gear),
                    bias_q=F(a.bias_q), bias_qd=F(a.bias_qd))
  ng = 0 if s.geom_bodyid is None else len(s.geom_bodyid)
  return rb.System(
      gravity=F(s.gravity), viscosity=F(s.viscosity), density=F(s.density), link=link, dof=dof, actuator=act, init_q=F(s.init_q),
      elasticity=jp.zeros(max(ng, 1)), vel_damping=0.0, ang_damping=0.0, baumgarte_erp=0.0, spring_mass_scale=0.0,
      spring_inertia_scale=0.0, joint_scale_ang=0.0, joint_scale_pos=0.0, collide_scale=0.0,
      enable_fluid=bool(s.enable_fluid), link_names=list(s.link_names), link_types=s.link_types, link_parents=tuple(s.link_parents),
      matrix_inv_iterations=int(s.matrix_inv_iterations), solver_iterations=int(s.solver_iterations),
      solver_maxls=int(s.solver_maxls), mj_model=None,
      nq=s.nq, nv=nv, nu=s.nu, opt=mjx.Option(timestep=float(np.float32(s.opt.timestep))),
      geom_bodyid=np.asarray(s.geom_bodyid, np.int32) if ng else np.zeros(0, np.int32),
      geom_pos=F(s.geom_pos) if ng else jp.zeros((0, 3)), geom_quat=F(s.geom_quat) if ng else jp.zeros((0, 4)))


LEAVES = {   # our flat State field -> accessor on the reference State
    'q': lambda t: t.q, 'qd': lambda t: t.qd, 'x_pos': lambda t: t.x.pos, 'x_rot': lambda t: t.x.rot,
    'xd_ang': lambda t: t.xd.ang, 'xd_vel': lambda t: t.xd.vel, 'root_com': lambda t: t.root_com,
    'cinr_pos': lambda t: t.cinr.transform.pos, 'cinr_rot': lambda t: t.cinr.transform.rot, 'cinr_i': lambda t: t.cinr.i,
    'cinr_mass': lambda t: t.cinr.mass, 'cd_ang': lambda t: t.cd.ang, 'cd_vel': lambda t: t.cd.vel,
    'cdof_ang': lambda t: t.cdof.ang, 'cdof_vel': lambda t: t.cdof.vel, 'cdofd_ang': lambda t: t.cdofd.ang,
    'cdofd_vel': lambda t: t.cdofd.vel, 'mass_mx': lambda t: t.mass_mx, 'mass_mx_inv': lambda t: t.mass_mx_inv,
    'con_jac': lambda t: t.con_jac, 'con_diag': lambda t: t.con_diag, 'con_aref': lambda t: t.con_aref,
    'qf_smooth': lambda t: t.qf_smooth, 'qf_constraint': lambda t: t.qf_constraint, 'qdd': lambda t: t.qdd,
}

# model -> (envs, steps, how far to drop the root towards the floor so that contacts are active)
CASES = {'ant': (3, 6, 0.0), 'humanoid': (2, 6, 0.0), 'halfcheetah': (2, 5, 0.35), 'hopper': (2, 5, 0.04),
         'walker2d': (2, 5, 0.05), 'humanoidstandup': (2, 4, 0.0), 'pusher': (2, 5, 0.0), 'triple_pendulum_motor': (2, 4, 0.0), 'inverted_pendulum': (2, 4, 0.0),
         'inverted_double_pendulum': (2, 4, 0.0), 'reacher': (2, 4, 0.0), 'swimmer': (2, 4, 0.0), 'two_trees': (2, 6, 0.0)}


def inputs(s, name, n, steps, drop, seed=0):
  rng = np.random.default_rng(seed)
  q = np.asarray(s.init_q, np.float64)[None] + rng.uniform(-0.1, 0.1, (n, s.nq))
  if name == 'ant':
    q[:, 2] = 0.45 + 0.1 * rng.uniform(size=n)        # feet touching
  if name == 'humanoidstandup':   # lying on the floor (init_q), small noise as the env's reset
    q = np.asarray(s.init_q, np.float64)[None] + rng.uniform(-0.01, 0.01, (n, s.nq))
  if name == 'humanoid':
    q = np.asarray(s.init_q, np.float64)[None] + rng.uniform(-0.01, 0.01, (n, s.nq))
    q[:, 2] = 1.29 + 0.02 * rng.uniform(size=n)       # feet touching (foot spheres reach z = 1.4 - 1.3)
  if name == 'two_trees':      # both spheres start slightly inside the floor; keep the root quaternions near unit
    q[:, 2] = 0.24; q[:, 9] = 0.29
  if drop:
    q[:, 1] -= drop
  qd = 0.1 * rng.standard_normal((n, s.nv))
  act = rng.uniform(-1, 1, (steps, n, s.nu))
  if name in ('humanoid', 'humanoidstandup'):
    act *= 0.4
  return q.astype(np.float32).astype(np.float64), qd.astype(np.float32).astype(np.float64), act.astype(np.float32).astype(np.float64)


def load(name):
  if name in ('ant', 'humanoid', 'halfcheetah', 'hopper', 'walker2d', 'humanoidstandup', 'pusher', 'inverted_pendulum', 'inverted_double_pendulum',
              'reacher', 'swimmer'):
    return envs_assets.load(name)
  if name == 'two_trees':      # our own synthetic model (tests/synthetic_models.py): several free roots
    from brax_b200.io import mjcf
    from tests.synthetic_models import TWO_TREES_XML
    return mjcf.loads(TWO_TREES_XML)
  from brax_b200.io import model_json
  return model_json.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.json'))


ENV_XML = {'ant.xml': 'ant', 'humanoid.xml': 'humanoid', 'half_cheetah.xml': 'halfcheetah', 'hopper.xml': 'hopper',
           'walker2d.xml': 'walker2d', 'humanoidstandup.xml': 'humanoidstandup', 'pusher.xml': 'pusher', 'inverted_pendulum.xml': 'inverted_pendulum',
           'inverted_double_pendulum.xml': 'inverted_double_pendulum', 'reacher.xml': 'reacher', 'swimmer.xml': 'swimmer'}


def _mjcf_load(path):
  """Stands in for brax.io.mjcf.load (MuJoCo's compiler): our compiled constants for that asset."""
  s = load(ENV_XML[os.path.basename(str(path))])
  mjx.PAIRS = s.contact_pairs() if s.geom_bodyid is not None and len(s.contact_pairs().geom1) else None
  return reference_system(s)


def env_golden(only=None):
  """The reference envs + wrappers, unmodified, on the stand-ins."""
  import jax
  mjcf_stub.load = _mjcf_load
  from brax.envs import ant as ref_ant, half_cheetah as ref_hc, humanoid as ref_hum   # the reference
  from brax.envs import hopper as ref_hop, walker2d as ref_walk                       # the reference
  from brax.envs import humanoidstandup as ref_hs, pusher as ref_pu
  from brax.envs import inverted_pendulum as ref_ip, inverted_double_pendulum as ref_idp, reacher as ref_re, swimmer as ref_sw
  from brax.envs.wrappers import training as ref_wrap                                 # the reference
  assert ref_wrap.__file__.startswith('/root/reference/')
  cases = {'ant': (ref_ant.Ant, 4, 7, 5), 'humanoid': (ref_hum.Humanoid, 3, 6, 4), 'halfcheetah': (ref_hc.Halfcheetah, 3, 5, 3),
           'hopper': (ref_hop.Hopper, 4, 6, 4), 'walker2d': (ref_walk.Walker2d, 3, 5, 4),
           'inverted_pendulum': (ref_ip.InvertedPendulum, 4, 6, 4), 'inverted_double_pendulum': (ref_idp.InvertedDoublePendulum, 4, 6, 4),
           'reacher': (ref_re.Reacher, 3, 6, 4), 'swimmer': (ref_sw.Swimmer, 3, 6, 4),
           'humanoidstandup': (ref_hs.HumanoidStandup, 2, 5, 3), 'pusher': (ref_pu.Pusher, 3, 6, 4)}
  for name, (cls, n, steps, ep_len) in cases.items():
    if only is not None and f'env_{name}' not in only:
      continue
    env = ref_wrap.wrap(cls(backend='generalized'), episode_length=ep_len, action_repeat=1)
    rng = np.random.default_rng(7)
    st = env.reset(jax.random.split(jax.random.PRNGKey(3), n))
    if name == 'pusher':   # arm lowered onto the table, the object under the wrist: both contact kinds active
      q = np.asarray(st.pipeline_state.q).copy()
      inner = env.env.env.env
      for e in range(n):
        q[e, :7] = 0.0
        q[e, 1] = (0.40, 0.44, 0.42)[e % 3]
        w = np.asarray(inner.pipeline_init(jp.array(q[e]), st.pipeline_state.qd[e]).x.pos[6])
        q[e, 7], q[e, 8] = w[1] + 0.07, w[0] - 0.40
      q = q.astype(np.float32).astype(np.float64)
      ps = jax.vmap(inner.pipeline_init)(jp.array(q), st.pipeline_state.qd)
      obs = jax.vmap(inner._get_obs)(ps)
      st = st.replace(pipeline_state=ps, obs=obs)
      st.info['first_pipeline_state'], st.info['first_obs'] = ps, obs
    if name in ('ant', 'hopper', 'inverted_pendulum', 'inverted_double_pendulum'):
      # one env starts unhealthy (Ant: z too high; Hopper: root angle out of range; the pendulums: pole(s) tipped over): terminates at once
      q = np.asarray(st.pipeline_state.q).copy()
      if name == 'ant':
        q[0, 2] = 1.3
      elif name == 'hopper':
        q[0, 2] = 0.5
      elif name == 'inverted_pendulum':
        q[0, 1] = 0.5
      else:
        q[0, 1] = 1.5
      inner = env.env.env.env     # AutoReset -> Episode -> Vmap -> the env
      ps = jax.vmap(inner.pipeline_init)(jp.array(q), st.pipeline_state.qd)
      obs = jax.vmap(inner._get_obs)(ps)
      st = st.replace(pipeline_state=ps, obs=obs)
      st.info['first_pipeline_state'], st.info['first_obs'] = ps, obs
    out = {'q0': np.asarray(st.pipeline_state.q), 'qd0': np.asarray(st.pipeline_state.qd), 'obs0': np.asarray(st.obs),
           'episode_length': np.array(ep_len)}
    acts = rng.uniform(-1, 1, (steps, n, env.action_size)).astype(np.float32).astype(np.float64)
    out['act'] = acts
    for k in range(steps):
      st = env.step(st, jp.array(acts[k]))
      rec = {'obs': st.obs, 'reward': st.reward, 'done': st.done, 'steps': st.info['steps'], 'truncation': st.info['truncation'],
             'q': st.pipeline_state.q, 'qd': st.pipeline_state.qd}
      rec.update({f'metric_{m}': v for m, v in st.metrics.items()})
      for f, get in LEAVES.items():
        rec[f'ps_{f}'] = get(st.pipeline_state)
      for kk, v in rec.items():
        out[f'step{k}_{kk}'] = np.asarray(v, np.float64)
    path = os.path.join(ROOT, 'tests', 'golden', f'ref_env_{name}.npz')
    np.savez_compressed(path, **out)
    dones = [float(out[f'step{k}_done'].sum()) for k in range(steps)]
    print(f'env {name}: {n} envs x {steps} env-steps (episode_length {ep_len}), done per step {dones}, wrote {path} '
          f'({os.path.getsize(path) // 1024} KB)')


def main():
  only = set(sys.argv[1].split(',')) if len(sys.argv) > 1 else None   # e.g. `swimmer,env_reacher`
  env_golden(only)
  for name, (n, steps, drop) in CASES.items():
    if only is not None and name not in only:
      continue
    s = load(name)
    mjx.PAIRS = s.contact_pairs() if s.geom_bodyid is not None and len(s.contact_pairs().geom1) else None
    rs = reference_system(s)
    q0, qd0, act = inputs(s, name, n, steps, drop)
    if name == 'pusher':   # arm lowered until the wrist capsules touch the table, the object pushed under the wrist
      for e in range(n):
        q0[e, :] = 0.0
        q0[e, 1] = (0.40, 0.44)[e % 2]
        w = np.asarray(ref_pipeline.init(rs, jp.array(q0[e]), jp.array(qd0[e])).x.pos[6])
        q0[e, 7], q0[e, 8] = w[1] + 0.07, w[0] - 0.40
      q0 = q0.astype(np.float32).astype(np.float64)
    out = {'q0': q0, 'qd0': qd0, 'act': act}
    active = 0
    for e in range(n):
      st = ref_pipeline.init(rs, jp.array(q0[e]), jp.array(qd0[e]))
      for f, get in LEAVES.items():
        out.setdefault(f'init_{f}', []).append(np.asarray(get(st), np.float64))
      for k in range(steps):
        st = ref_pipeline.step(rs, st, jp.array(act[k, e]))
        for f, get in LEAVES.items():
          out.setdefault(f'step{k}_{f}', []).append(np.asarray(get(st), np.float64))
        active += int((np.asarray(st.con_diag) != 0).sum())
    out = {k: (np.stack(v) if isinstance(v, list) else v) for k, v in out.items()}
    path = os.path.join(ROOT, 'tests', 'golden', f'ref_{name}.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: {n} envs x {steps} steps, active constraint rows seen {active}, wrote {path} ({os.path.getsize(path) // 1024} KB)')


if __name__ == '__main__':
  main()
