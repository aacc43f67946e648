#!/usr/bin/env python
"""bench.py -- env-steps/sec of the generalized physics step (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one env-step (n_frames = 5 physics substeps, one action) over the
whole env batch of this rank.  Prints ONE JSON line (rank 0).  Under torchrun
each rank owns a contiguous shard of the env batch on its own GPU; there is no
collective on the step path (scaling = weak: per-GPU envs fixed).

Workloads (BASELINE.json `configs`):
  humanoid_8192  configs[1]  Humanoid, 8192 envs/GPU            (default at N=1)
  ant_1m         configs[2]  Ant, 1,048,576 envs/GPU
  humanoid_512k  configs[3]  Humanoid, 524,288 envs/GPU (4 M over 8 GPUs), randomised falls
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

WORKLOADS = {
    'humanoid_8192': ('humanoid', 8192),
    'ant_1m': ('ant', 1 << 20),
    'humanoid_512k': ('humanoid_falls', 1 << 19),
    'ant_1024': ('ant', 1024),
}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # non-tensor FMA peak at max clock


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
  """Samples SM clocks / throttle reasons of one GPU during the timed region: NVML in a thread
  (every 5 ms, no start-up latency), nvidia-smi as a fallback."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.gpu, self.rows, self.proc, self.nvml = gpu_index, [], None, None
    self.sm, self.reasons, self.max_mhz, self._stop = [], set(), None, False

  def _nvml_loop(self):
    n = self.nvml
    names = {getattr(n, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
             getattr(n, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
             getattr(n, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
             getattr(n, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap'}
    get_reasons = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
    while not self._stop:
      try:
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        bits = int(get_reasons(self.h))
        for bit, nme in names.items():
          if bits & bit:
            self.reasons.add(nme)
      except Exception:
        pass
      time.sleep(0.005)

  def start(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
      vis = os.environ.get('CUDA_VISIBLE_DEVICES')
      idx = self.gpu
      if vis and all(t.strip().isdigit() for t in vis.split(',')):
        idx = int(vis.split(',')[self.gpu])
      self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
      self.nvml = pynvml
      self.t = threading.Thread(target=self._nvml_loop, daemon=True)
      self.t.start()
      return
    except Exception:
      self.nvml = None
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if self.nvml is not None:
      self._stop = True
      self.t.join(timeout=1)
      sm = sorted(self.sm)
      return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'samples': len(sm),
              'reasons': sorted(self.reasons), 'source': 'nvml'}
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      try:
        sm.append(float(r[1])); mx.append(float(r[2]))
        for nme, v in zip(names, r[5:9]):
          if v.lower().startswith('active'):
            reasons.add(nme)
      except Exception:
        pass
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'samples': len(sm), 'reasons': sorted(reasons), 'source': 'nvidia-smi'}


def cpu_reference(model, steps, warmup, seconds_target=None, seed=0):
  """The reference arm AND the cpu_baseline leg: the C oracle (a port of the reference's algorithm, NOT JAX: jax /
  jaxopt / mujoco are absent from the image and from the GPU box, profiles/r02_probe_reference.log) on the host
  cores, OpenMP over envs with an EXPLICIT thread count (launchers export OMP_NUM_THREADS=1), on a bounded sample
  of the workload: 32 envs per thread.  With `seconds_target` the number of timed env-steps is chosen so that the
  run takes about that long.  Returns a dict."""
  import numpy as np
  from brax_b200 import workloads
  from oracle import oracle as O
  cores = os.cpu_count() or 1
  n = 32 * cores
  sys_, q, qd = workloads.reset(model, 0, n, seed, 'cpu')
  o = O.Oracle(sys_, np.float32, threads=cores)
  st = o.init(q.numpy(), qd.numpy())
  nf = workloads.N_FRAMES[model]
  t0 = time.perf_counter()
  for w in range(max(warmup, 1)):
    o.step(st, workloads.action(model, 0, n, seed, w, 'cpu').numpy(), nf)
  per_step = (time.perf_counter() - t0) / max(warmup, 1)
  if seconds_target is not None:
    steps = max(1, min(400, int(seconds_target / max(per_step, 1e-6))))
  acts = [workloads.action(model, 0, n, seed, 100 + k, 'cpu').numpy() for k in range(min(steps, 16))]
  t0 = time.perf_counter()
  for k in range(steps):
    o.step(st, acts[k % len(acts)], nf)
  dt = time.perf_counter() - t0
  return {'value': n * steps / dt, 'ms_per_step': 1e3 * dt / steps, 'cores': o.threads, 'n_env': n, 'steps': steps, 'nf': nf,
          'sample': f'{model}: {n} envs x {steps} env-steps, C oracle (-O2, no FMA contraction), OpenMP {o.threads} threads'}


def run_reference(args, rank, world):
  """--impl reference: the reference's CPU implementation of the path (the oracle port: see cpu_reference)."""
  if rank != 0:
    return
  model, n_env = WORKLOADS[args.workload]
  r = cpu_reference(model, args.steps, args.warmup)
  line = {
      'impl': 'reference', 'metric': 'env-steps/sec', 'value': r['value'], 'unit': 'env-steps/s', 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': args.workload, 'model': model, 'envs_per_gpu': n_env, 'n_frames': r['nf'],
                 'note': 'CPU restatement (C oracle port), not JAX: jax/jaxopt/mujoco unavailable offline; host threads set explicitly'},
      'cpu_baseline': {'value': r['value'], 'unit': 'env-steps/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']},
      'e2e': {'value': r['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=100)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--workload', default='humanoid_8192', choices=sorted(WORKLOADS))
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--envs', type=int, default=0, help='override envs per GPU')
  ap.add_argument('--minv', default='newton_schulz', choices=['newton_schulz', 'cholesky'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-extra', action='store_true', help='skip the secondary workload lines')
  ap.add_argument('--lean', action='store_true', help='BXG_STEP_LEAN state I/O (q, qd, x, xd, mass_mx_inv only) for the main workload')
  ap.add_argument('--no-ppo', action='store_true', help='skip the end-to-end PPO line (config 5)')
  ap.add_argument('--ppo-timesteps', type=int, default=6_000_000, help='env-steps per GPU of the PPO line')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))

  if args.impl == 'reference':
    run_reference(args, rank, world)
    return

  import torch
  import torch.distributed as dist
  from brax_b200 import native, workloads

  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)

  def barrier():
    if world > 1:
      dist.barrier(device_ids=[local_rank])
    torch.cuda.synchronize()

  def max_over_ranks(x):
    if world > 1:
      t = torch.tensor([x], dtype=torch.float64, device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      return float(t.item())
    return x

  minv = native.MINV_NEWTON_SCHULZ if args.minv == 'newton_schulz' else native.MINV_CHOLESKY
  hbm_peak, peak_src = peaks()
  flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

  def measure(workload, steps, warmup, with_clocks, minv=minv, lean=False):
    model, n_env = WORKLOADS[workload]
    if args.envs and workload == args.workload:
      n_env = args.envs
    nf = workloads.N_FRAMES[model]
    begin = rank * n_env  # weak scaling: every rank has n_env envs; ids are global
    sys_, q, qd = workloads.reset(model, begin, n_env, 0, dev)
    nm = native.model_for(sys_, local_rank, minv)
    a, b = nm.init(q, qd), nm.alloc(n_env, lean)
    if lean:   # the derived State leaves are neither read nor written in this mode: drop them
      a = {k: a[k] for k in native.LEAN_FIELDS}
    acts = [workloads.action(model, begin, n_env, 0, k, dev) for k in range(min(steps + warmup, 8))]
    state_bytes = sum(t.numel() * 4 for t in a.values())
    flush = state_bytes < (512 << 20)
    for w in range(warmup):
      nm.step(a, acts[w % len(acts)], nf, out=b, lean=lean); a, b = b, a
    sampler = ClockSampler(local_rank) if with_clocks else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    if sampler:
      sampler.start()
    l0 = native.launch_count()
    t0 = time.perf_counter()
    for k in range(steps):
      if flush:
        flush_buf.fill_(k & 0xff)
      ev[k][0].record()
      nm.step(a, acts[(warmup + k) % len(acts)], nf, out=b, lean=lean)
      ev[k][1].record()
      a, b = b, a
    barrier()
    wall = time.perf_counter() - t0
    launches = native.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    kern_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    dev_ms_total = ev[0][0].elapsed_time(ev[-1][1])  # first start -> last end, on-device
    dev_ms_total = max_over_ranks(dev_ms_total)
    kern_avg_ms = max_over_ranks(sum(kern_ms) / len(kern_ms))
    # the reference algorithm itself diverges for a few envs under long random-action rollouts without
    # termination (profiles/r01_rollout_*.json): report them, do not hide them
    nonfinite = int((~torch.isfinite(a['q']).all(dim=1)).sum())
    assert nonfinite <= max(1, n_env // 20), f'{nonfinite} of {n_env} envs non-finite'
    total_envs = n_env * world
    return {
        # the timed quantity is the K steps themselves (sum of the per-step CUDA-event intervals, max over
        # ranks); the L2 flush that separates them is measurement scaffolding and is reported beside it
        'model': model, 'n_env': n_env, 'nf': nf, 'value': total_envs / (kern_avg_ms * 1e-3),
        'ms_per_step': kern_avg_ms, 'kern_avg_ms': kern_avg_ms, 'wall_ms_per_step': 1e3 * wall / steps,
        'ms_per_step_incl_flush': dev_ms_total / steps,
        'launches': launches, 'clocks': clocks, 'flush': flush, 'nm': nm, 'minv': minv, 'lean': lean, 'state_bytes': state_bytes, 'state': a, 'spare': b, 'acts': acts, 'nonfinite': nonfinite,
        'begin': begin,
    }

  def measure_e2e(r, steps):
    """Same metric through the public call with HOST buffers: per step, H2D copy of
    the actions from pinned memory, the step, D2H read of q and qd."""
    nm, a, b = r['nm'], r['state'], r['spare']
    n_env, nf, model = r['n_env'], r['nf'], r['model']
    sys_ = nm.sys
    h_act = [t.cpu().pin_memory() for t in r['acts']]
    d_act = torch.empty_like(r['acts'][0])
    h_q = torch.empty((n_env, sys_.nq), dtype=torch.float32).pin_memory()
    h_qd = torch.empty((n_env, sys_.nv), dtype=torch.float32).pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
      d_act.copy_(h_act[k % len(h_act)], non_blocking=True)
      nm.step(a, d_act, nf, out=b, lean=r['lean'])
      a, b = b, a
      h_q.copy_(a['q'], non_blocking=True); h_qd.copy_(a['qd'], non_blocking=True)
      torch.cuda.current_stream().synchronize()  # the caller consumes the result every step
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    return {'value': n_env * world * steps / (ms * 1e-3), 'unit': 'env-steps/s',
            'h2d_bytes_per_step': n_env * sys_.nu * 4, 'd2h_bytes_per_step': n_env * (sys_.nq + sys_.nv) * 4}

  def flops_per_env_step(sysm, nf):
    # algorithmic FLOPs per env-step at the reference iteration counts: Newton-Schulz
    # (2*iters + 1 products of 2 nv^3), A = J Minv J^T, ~6.6 matvecs per solver iteration, O(L) terms
    nv_, nc_ = sysm.nv, native.num_constraints(sysm)
    per_sub = ((2 * sysm.matrix_inv_iterations + 1) * 2 * nv_ ** 3 + 2 * nc_ * nv_ ** 2 + 2 * nc_ ** 2 * nv_
               + sysm.solver_iterations * 6.6 * 2 * nc_ ** 2 + 20e3)
    return per_sub * nf

  def static_traffic(workload, n_env):
    """DRAM bytes per launch from the committed ncu capture of this workload (profiles/traffic.json): a constant of
    the kernel's I/O layout (full State in, full State out), NOT measured in this run."""
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
      with open(tp) as f:
        tj = json.load(f)
      if workload in tj:
        return tj[workload]['dram_bytes_per_env_step'] * n_env
    return None

  def summarize(workload, r, e2e):
    model = r['model']
    tkey = workload + ('_lean' if r['lean'] else '')
    achieved = workloads.ALGO_BYTES[model] * r['n_env'] / (r['kern_avg_ms'] * 1e-3) / 1e9     # per launch = per rank per step
    fl = flops_per_env_step(r['nm'].sys, r['nf'])
    tf = fl * r['n_env'] / (r['kern_avg_ms'] * 1e-3) / 1e12
    plan = native.plan(r['nm'].sys, r['minv'])
    shape = r['nm'].launch_shape(r['n_env'])
    return {
        'value': r['value'], 'unit': 'env-steps/s', 'ms_per_step': r['ms_per_step'], 'e2e': e2e,
        'launch': {'variant': plan['variant'], 'lanes_per_env': plan['lanes_per_env'], 'grid': shape['grid'],
                   'threads_per_cta': shape['threads_per_cta'], 'envs_per_cta': shape['envs_per_cta'],
                   'smem_bytes_per_cta': shape['smem_bytes_per_cta'], 'envs_per_cta_max': plan['envs_per_cta']},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
                     'traffic': static_traffic(tkey, r['n_env']),
                     'traffic_source': 'static: ncu dram__bytes of the committed capture (profiles/traffic.json), not measured in this run',
                     'peak_source': peak_src, 'algorithmic_bytes_per_env_step': workloads.ALGO_BYTES[model],
                     'note': 'the step is bound on-chip (shared-memory operand delivery and FMA issue of the Newton-Schulz products, '
                             'then latency), not by HBM (SURVEY 8d); see profiles/ for pipe utilisation'},
        'fp32': {'achieved': tf, 'peak': FP32_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': tf / FP32_PEAK_TFLOPS, 'flops_per_env_step': fl,
                 'note': 'theoretical non-tensor FP32 FMA peak (148 SMs x 128 lanes x 2 x 1.965 GHz), not in MEASURED_PEAKS.json; '
                         'the dense work is fp32 by the parity requirement'},
        'l2': 'flushed between steps (256 MiB write, outside the per-step event intervals)' if r['flush'] else 'state >> L2, no flush',
        'nonfinite_envs': r['nonfinite'],
        'state_io': 'lean (q, qd, x, xd, mass_mx_inv; BXG_STEP_LEAN)' if r['lean'] else 'full generalized.State (25 leaves, drop-in)',
        'state_bytes_per_env': r['state_bytes'] // r['n_env'],
    }

  r = measure(args.workload, args.steps, args.warmup, with_clocks=True, lean=args.lean)
  e2e = measure_e2e(r, args.steps)
  model = r['model']
  sm = summarize(args.workload, r, e2e)
  line = {
      'metric': 'env-steps/sec', 'value': r['value'], 'unit': 'env-steps/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': args.workload, 'model': model, 'envs_per_gpu': r['n_env'], 'n_frames': r['nf'],
                 'minv': args.minv, 'parallelism': f'env-shard x{world}, no collective',
                 'launch': sm['launch'], 'l2': sm['l2'], 'state_io': sm['state_io']},
      'roofline': sm['roofline'], 'fp32': sm['fp32'],
      'e2e': e2e, 'gpu_launches': r['launches'], 'clocks': r['clocks'],
      'kernel_ms': r['kern_avg_ms'], 'wall_ms_per_step': r['wall_ms_per_step'],
      'ms_per_step_incl_l2_flush': r['ms_per_step_incl_flush'], 'nonfinite_envs': r['nonfinite'],
  }

  if not args.no_extra:
    # The metric's other configurations, at EVERY N (BASELINE configs 3 and 4: Ant 1 M envs per GPU; Humanoid 512 k
    # envs per GPU = 4 M over 8 GPUs, randomised falls), each with its own e2e leg and both roofline fractions; then
    # the exact-inverse mode on the default workload (north_star's per-env Cholesky solve: NOT the reference's
    # Newton-Schulz numerics); then config 5 (end-to-end PPO on Ant).
    extra = []
    del r['state'], r['spare'], r['nm'], r['acts']
    for wl, st_, mv, ln in (('ant_1m', 5, minv, False), ('humanoid_512k', 5, minv, False), ('ant_1m', 5, minv, True),
                            (args.workload, 10, minv, True), (args.workload, 10, native.MINV_CHOLESKY, False)):
      if wl == args.workload and mv == minv and ln == args.lean:
        continue
      try:
        torch.cuda.empty_cache()
        x = measure(wl, st_, 3, with_clocks=False, minv=mv, lean=ln)
        xe = measure_e2e(x, st_)
        ent = {'workload': wl, 'model': x['model'], 'envs_per_gpu': x['n_env'], 'n_gpus': world, 'steps': st_,
               'minv': 'cholesky' if mv == native.MINV_CHOLESKY else 'newton_schulz'}
        ent.update(summarize(wl, x, xe))
        extra.append(ent)
        del x
      except Exception as ex:  # report, do not hide
        extra.append({'workload': wl, 'error': repr(ex)})
    if not args.no_ppo:
      try:
        torch.cuda.empty_cache()
        from brax_b200.training import ppo
        # notebooks/training.ipynb:250 hyper-parameters; a bounded number of env-steps
        _, pm = ppo.train('ant', num_envs=4096, episode_length=1000, num_timesteps=args.ppo_timesteps * world, unroll_length=5,
                          batch_size=2048, num_minibatches=32, num_update_epochs=4, reward_scaling=10.0, entropy_cost=1e-2,
                          discounting=0.97, learning_rate=3e-4, seed=1, device=dev)
        extra.append({'workload': 'ppo_ant', 'config': 'BASELINE configs[4]: end-to-end PPO on Ant, 4096 envs per GPU, unroll 5, 32 minibatches x 2048, '
                      '4 epochs; env-steps/s include policy inference and the learner; gradients and normaliser statistics all-reduced over NCCL',
                      'n_gpus': world, 'value': max_over_ranks(-pm['sps_steady']) * -1.0, 'unit': 'env-steps/s',
                      'value_incl_graph_capture': pm['sps'], 'env_steps': pm['env_steps'], 'iterations': pm['iterations'],
                      'episode_reward': pm['episode_reward']})
      except Exception as ex:
        extra.append({'workload': 'ppo_ant', 'error': repr(ex)})
    line['other_workloads'] = extra

  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    c = cpu_reference(model, 0, 2, seconds_target=12.0)
    line['cpu_baseline'] = {'value': c['value'], 'unit': 'env-steps/s', 'cores': c['cores'], 'kind': 'port', 'sample': c['sample']}

  if rank == 0:
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
