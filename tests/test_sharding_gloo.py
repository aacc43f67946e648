"""N > 1 host logic on CPU: world_size-2 gloo processes each take their shard of
the env batch; inputs are keyed by the global env id so the union of the shards
equals the single-process batch bit for bit.  The oracle stands in for the
physics here (this is a test of the sharding plumbing, not of the kernel)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
  return p


def _worker(rank, world, port, n_total, out_dir):
  os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from brax_b200 import sharding, workloads
  from oracle import oracle as O
  begin, end = sharding.shard_range(n_total, rank, world)
  sys_, q, qd = workloads.reset('ant', begin, end - begin, 0, 'cpu')
  o = O.Oracle(sys_, threads=1)
  st = o.init(q.numpy(), qd.numpy())
  for k in range(2):
    o.step(st, workloads.action('ant', begin, end - begin, 0, k, 'cpu').numpy(), 5)
  # the only cross-rank traffic: timing-style max-reduce and a checksum gather
  t = torch.tensor([float(rank + 1)])
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  assert t.item() == world
  cs = torch.tensor([float(st['q'].astype(np.float64).sum())], dtype=torch.float64)
  gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
  dist.all_gather(gathered, cs)
  np.save(os.path.join(out_dir, f'q_{rank}.npy'), st['q'])
  if rank == 0:
    np.save(os.path.join(out_dir, 'checksums.npy'), np.array([g.item() for g in gathered]))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_shards_equal_single_batch(tmp_path):
  from brax_b200 import sharding, workloads
  from oracle import oracle as O
  n_total, world = 13, 2            # ragged split: 7 + 6
  port = _free_port()
  mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
  sys_, q, qd = workloads.reset('ant', 0, n_total, 0, 'cpu')
  o = O.Oracle(sys_, threads=1)
  st = o.init(q.numpy(), qd.numpy())
  for k in range(2):
    o.step(st, workloads.action('ant', 0, n_total, 0, k, 'cpu').numpy(), 5)
  parts = np.concatenate([np.load(tmp_path / f'q_{r}.npy') for r in range(world)])
  assert np.array_equal(parts, st['q'])
  cs = np.load(tmp_path / 'checksums.npy')
  assert abs(cs.sum() - st['q'].astype(np.float64).sum()) < 1e-9


def test_shard_ranges_cover_and_partition():
  from brax_b200 import sharding
  for n in (0, 1, 7, 8192, 1 << 20):
    for w in (1, 2, 4, 8):
      r = [sharding.shard_range(n, k, w) for k in range(w)]
      assert r[0][0] == 0 and r[-1][1] == n
      assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
      sizes = [b - a for a, b in r]
      assert max(sizes) - min(sizes) <= 1


def test_inputs_depend_only_on_global_env_id():
  from brax_b200 import sharding
  full = sharding.uniform(0, 64, 5, seed=3, stream=7)
  for begin, n in ((0, 10), (10, 30), (40, 24)):
    assert torch.equal(sharding.uniform(begin, n, 5, seed=3, stream=7), full[begin:begin + n])
  a = sharding.normal(0, 100000, 2, 1, 1)
  assert abs(a.mean().item()) < 0.02 and abs(a.std().item() - 1) < 0.02
