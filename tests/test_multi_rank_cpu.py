"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: envs shard by contiguous index blocks with no
data-path collective, so (a) the shards partition the env range, (b) stepping the shards independently and
concatenating equals stepping the whole batch (checked with the oracle, which is what runs on CPU), and
(c) the barrier / max-over-ranks timing helpers of bench.py work across processes."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from gridworld_b200 import dist as D
    from oracle import oracle as O
    r, w = D.init('gloo')
    assert (r, w) == (rank, world)
    n, steps = 96, 40
    rng = np.random.RandomState(0)  # same stream on every rank: the global problem
    T = np.where(rng.random_sample((n, 9, 11, 11)) < 0.1, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8)
    T[:, 4:] = 0
    acts = rng.randint(0, 18, (steps, n)).astype(np.int32)
    lo, hi = D.shard_range(n, rank, world)
    b = O.OracleBatch(hi - lo, 'walking', max_steps=15)
    b.set_tasks(T[lo:hi])
    D.barrier()
    rew = np.zeros((steps, hi - lo))
    for t in range(steps):
        rew[t], _ = b.step_walking(acts[t, lo:hi], auto_reset=True)
    D.barrier()
    # task generator: a shard passes env_offset = lo, so its envs draw what they draw in the unsharded batch
    g = O.OracleBatch(hi - lo, 'walking', max_steps=7)
    g.set_pool(T[:13], None, seed=99, env_offset=lo)
    g.reset()
    grew = np.zeros((steps, hi - lo))
    for t in range(steps):
        grew[t], _ = g.step_walking(acts[t, lo:hi], auto_reset=True)
    np.savez(os.path.join(tmpdir, f'shard{rank}.npz'), lo=lo, hi=hi, rew=rew, grid=b.state()['grid'],
             pos=b.state()['pos'], grew=grew, gids=g.task_ids.copy(), ggrid=g.state()['grid'])
    # timing helpers: the slowest rank defines the time, the env counts add up
    assert D.max_over_ranks(1.0 + rank) == float(world)
    assert D.sum_over_ranks(hi - lo) == float(n)
    D.shutdown()


def test_two_ranks_shard_and_concatenate(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from gridworld_b200 import dist as D
    n, steps = 96, 40
    rng = np.random.RandomState(0)
    T = np.where(rng.random_sample((n, 9, 11, 11)) < 0.1, rng.randint(1, 7, (n, 9, 11, 11)), 0).astype(np.int8)
    T[:, 4:] = 0
    acts = rng.randint(0, 18, (steps, n)).astype(np.int32)
    full = O.OracleBatch(n, 'walking', max_steps=15)
    full.set_tasks(T)
    rew = np.zeros((steps, n))
    for t in range(steps):
        rew[t], _ = full.step_walking(acts[t], auto_reset=True)
    shards = [np.load(tmp_path / f'shard{r}.npz') for r in range(world)]
    assert [int(s['lo']) for s in shards] == [0, 48] and [int(s['hi']) for s in shards] == [48, 96]
    assert np.array_equal(np.concatenate([s['rew'] for s in shards], axis=1), rew)
    assert np.array_equal(np.concatenate([s['grid'] for s in shards]), full.state()['grid'])
    assert np.array_equal(np.concatenate([s['pos'] for s in shards]), full.state()['pos'])
    gfull = O.OracleBatch(n, 'walking', max_steps=7)
    gfull.set_pool(T[:13], None, seed=99, env_offset=0)
    gfull.reset()
    grew = np.zeros((steps, n))
    for t in range(steps):
        grew[t], _ = gfull.step_walking(acts[t], auto_reset=True)
    assert np.array_equal(np.concatenate([s['grew'] for s in shards], axis=1), grew)
    assert np.array_equal(np.concatenate([s['gids'] for s in shards]), gfull.task_ids)
    assert np.array_equal(np.concatenate([s['ggrid'] for s in shards]), gfull.state()['grid'])
    # uneven split
    assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_single_rank_helpers():
    sys.path.insert(0, ROOT)
    from gridworld_b200 import dist as D
    assert D.shard_range(65536, 0, 1) == (0, 65536)
    assert D.max_over_ranks(3.5) == 3.5 and D.sum_over_ranks(2) == 2.0
