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
