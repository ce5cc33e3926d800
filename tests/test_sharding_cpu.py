"""Host-side multi-GPU logic on CPU: world_size 2 over gloo (weight broadcast, disjoint game ids, metric reduction)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _pkg  # noqa: F401  (spawned workers re-import this module without conftest)
import alphazero_chess_b200 as az
from alphazero_chess_b200 import sharding


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = az.weight_sizes()
    offs = sharding.weight_offsets(sizes)
    flat = torch.zeros(int(offs[-1]), dtype=torch.float32)
    if rank == 0:
        flat.copy_(torch.from_numpy(sharding.flatten_weights(az.random_weights(seed=42, sizes=sizes, names=az.weight_names()))))
    sharding.broadcast_weights(flat, dist, src=0)
    # work counters: rank r reports (r+1) * 1000 simulations in (r+1) * 10 ms
    sums, maxes = sharding.reduce_metrics([1000.0 * (rank + 1), 7.0], [10.0 * (rank + 1)], dist)
    np.save(os.path.join(out_dir, f"w{rank}.npy"), flat.numpy())
    np.save(os.path.join(out_dir, f"m{rank}.npy"), np.concatenate([sums, maxes, [sharding.first_game_id(rank)]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    w0, w1 = np.load(tmp_path / "w0.npy"), np.load(tmp_path / "w1.npy")
    ref = sharding.flatten_weights(az.random_weights(seed=42))
    assert np.array_equal(w0, ref) and np.array_equal(w1, ref)
    m0, m1 = np.load(tmp_path / "m0.npy"), np.load(tmp_path / "m1.npy")
    assert np.allclose(m0[:3], [3000.0, 14.0, 20.0]) and np.allclose(m1[:3], m0[:3])
    assert m0[3] == 0 and m1[3] == float(1 << 40)


def test_split_weights_round_trip():
    sizes = az.weight_sizes()
    arrays = az.random_weights(seed=3)
    parts = sharding.split_weights(sharding.flatten_weights(arrays), sizes)
    assert all(np.array_equal(a, b) for a, b in zip(arrays, parts))
    ids = [sharding.first_game_id(r) for r in range(8)]
    assert len(set(ids)) == 8 and all(b - a == sharding.GAME_ID_STRIDE for a, b in zip(ids, ids[1:]))


def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rounds = []
    # round 0: both ranks contribute (3 and 5 records), round 1: only rank 1, round 2: nobody
    for rnd, counts in enumerate(([3, 5], [0, 4], [0, 0])):
        mine = np.zeros(counts[rank], az.SAMPLE_DTYPE)
        mine["game_id"] = sharding.first_game_id(rank) + np.arange(counts[rank])
        mine["ply"] = rnd
        mine["final_value"] = rank + 0.5
        got = sharding.gather_samples(mine, dist)
        rounds.append(got)
        done = sharding.all_done(rank == 1 or rnd >= 1, dist)
        assert done == (rnd >= 1)
    np.save(os.path.join(out_dir, f"g{rank}.npy"), np.concatenate(rounds).view(np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_gather_two_rank_gloo(tmp_path):
    """Replay samples reach rank 0 in rank order with every byte intact; other ranks receive nothing."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0 = np.load(tmp_path / "g0.npy").view(az.SAMPLE_DTYPE)
    g1 = np.load(tmp_path / "g1.npy").view(az.SAMPLE_DTYPE)
    assert len(g1) == 0 and len(g0) == 12
    stride = sharding.GAME_ID_STRIDE
    assert list(g0["game_id"]) == [0, 1, 2, stride, stride + 1, stride + 2, stride + 3, stride + 4, stride, stride + 1, stride + 2, stride + 3]
    assert list(g0["ply"]) == [0] * 8 + [1] * 4
    assert list(g0["final_value"][:3]) == [0.5] * 3 and list(g0["final_value"][3:]) == [1.5] * 9
