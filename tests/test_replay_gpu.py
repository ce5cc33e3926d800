"""memory.rs ReplayBuffer on the GPU against the oracle restatement: identical entries (bit-exact running means),
identical unique counts and FIFO eviction, and the sampling contract."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import orc

pytestmark = pytest.mark.gpu

SIMS = 16


@pytest.fixture(scope="module")
def played():
    """Self-play records with many repeated early positions (synthetic evaluator, 48 games)."""
    e = az.Engine(max_games=48, num_simulations=SIMS, seed=3)
    e.set_evaluator_stub(1, 8)
    e.selfplay_begin(48)
    chunks = []
    for _ in range(300):
        st = e.selfplay_step(64)
        if st.pending_samples:
            chunks.append(e.selfplay_drain())
        if st.games_finished >= 60:
            break
    samples = np.concatenate(chunks)
    assert len(samples) > 1500
    yield e, samples
    e.close()


def _oracle_fill(rep, samples):
    new_unique = 0
    for s in samples:
        new_unique += rep.add(s["position"], az.improved_policy(s, SIMS), float(s["final_value"]))
    return new_unique


@pytest.mark.parametrize("capacity", [100_000, 300])
def test_add_matches_oracle(played, capacity):
    e, samples = played
    gpu = az.ReplayBuffer(e, capacity=capacity, max_batch=64)
    ref = orc.Replay(capacity)
    # several batches, as run_all_episodes would deliver them per generation
    total_new = 0
    for part in np.array_split(samples, 5):
        got = gpu.add(part)
        want = _oracle_fill(ref, part)
        assert got == want
        total_new += got
        assert len(gpu) == len(ref)
    assert len(gpu) == min(capacity, total_new)
    assert total_new < len(samples)  # the games share their openings: de-duplication happened
    # every position ever seen: same entry (or same absence after eviction)
    seen = {}
    for s in samples:
        seen[s["position"].tobytes()] = s["position"]
    repeats = 0
    for pos in list(seen.values())[:1200]:
        gp, gv, gn = gpu.get(pos)
        rp, rv, rn = ref.get(pos)
        assert gn == rn
        if rn:
            assert gv == np.float32(rv) and np.array_equal(gp, rp)
            repeats += rn > 1
    if capacity > len(samples):
        assert repeats > 10     # shared openings really were merged by running means
    else:
        assert len(gpu) == capacity  # FIFO eviction kept exactly the newest `capacity` unique positions
    gpu.close()


def test_add_pending_consumes_device_samples(played):
    e, _ = played
    e.selfplay_begin(48, first_game_id=10_000)
    gpu = az.ReplayBuffer(e, capacity=50_000, max_batch=64)
    ref = orc.Replay(50_000)
    added = 0
    for _ in range(200):
        st = e.selfplay_step(64)
        if st.pending_samples and added == 0:
            # same samples to both: read them (drain would reset the queue, so peek through a second engine-side add)
            n, nu = gpu.add_pending()
            assert n == st.pending_samples and nu > 0 and len(gpu) == nu
            added = n
            st2 = e.selfplay_step(0)
            assert st2.pending_samples == 0
            break
    assert added > 0
    gpu.close()
    del ref


def test_sample_contract(played):
    e, samples = played
    gpu = az.ReplayBuffer(e, capacity=100_000, max_batch=512)
    assert gpu.sample(512)[0].shape[0] == 0           # empty buffer -> empty batch (memory.rs:80-82)
    gpu.add(samples[:40])
    n_unique = len(gpu)
    planes, policy, value = gpu.sample(512, seed=1)   # fewer entries than the batch: all of them, once each
    assert planes.shape[0] == n_unique
    assert len({p.tobytes() for p in planes}) == n_unique
    gpu.add(samples[40:])
    planes, policy, value = gpu.sample(512, seed=2)
    assert planes.shape == (512, 19, 8, 8) and policy.shape == (512, 4096)
    assert len({p.tobytes() + q.tobytes() for p, q in zip(planes, policy)}) == 512   # without replacement
    assert np.allclose(policy.sum(1), 1.0, atol=1e-5)
    # rows are real entries: the planes are to_tensor of a stored position and the policy/value its running means
    by_planes = {}
    for s in samples:
        by_planes.setdefault(orc.to_tensor(s["position"]).tobytes(), s["position"])
    for k in range(0, 512, 37):
        pos = by_planes[planes[k].tobytes()]
        gp, gv, gn = gpu.get(pos)
        assert gn >= 1 and np.array_equal(gp, policy[k]) and gv == value[k]
    p2, _, _ = gpu.sample(512, seed=3)
    assert p2.tobytes() != planes.tobytes()
    # the device-resident path (az_replay_sample_dev into CUDA tensors) returns the very same batch
    tp, tpi, tv = gpu.sample_torch(512, seed=2)
    assert tp.is_cuda and np.array_equal(tp.cpu().numpy(), planes) and np.array_equal(tpi.cpu().numpy(), policy)
    assert np.array_equal(tv.cpu().numpy(), value)
    gpu.close()


@pytest.mark.parametrize("capacity", [100_000, 300])
def test_save_load_round_trip(played, capacity, tmp_path):
    """ReplayBuffer::save / load (memory.rs:100-115): a reloaded buffer holds the same entries in the same FIFO order, so it
    samples identically and evicts identically when more steps arrive."""
    e, samples = played
    a = az.ReplayBuffer(e, capacity=capacity, max_batch=64)
    first, rest = samples[: len(samples) * 2 // 3], samples[len(samples) * 2 // 3:]
    a.add(first)
    path = tmp_path / "replay_buffer"
    a.save(path)
    b = az.ReplayBuffer(e, capacity=capacity, max_batch=64)
    assert b.load(path) == len(a) == len(b)
    pa, pb = a.export(0, 64), b.export(0, 64)
    for x, y in zip(pa, pb):
        assert x.tobytes() == y.tobytes()
    for seed in (0, 9):
        for x, y in zip(a.sample(48, seed=seed), b.sample(48, seed=seed)):
            assert np.array_equal(x, y)
    assert a.add(rest) == b.add(rest)          # same de-duplication and the same evictions afterwards
    assert len(a) == len(b)
    n = len(a)
    for start in (0, max(0, n - 64)):
        for x, y in zip(a.export(start, 64), b.export(start, 64)):
            assert x.tobytes() == y.tobytes()
    # the file itself: stored FENs carry the pseudo-legal en-passant square only, visit counts >= 1
    from alphazero_chess_b200 import replay_io
    pos, policy, value, visits = replay_io.read_file(path)
    assert visits.min() >= 1 and np.allclose(policy.sum(1), 1.0, atol=1e-3)
    a.close()
    b.close()


@pytest.mark.parametrize("env", [{"AZ_REPLAY_FORCE_SLOW": "1"}, {"AZ_REPLAY_WINDOW": "1"}, {"AZ_REPLAY_WINDOW": "5"}])
def test_add_paths_agree_with_oracle(env):
    """Phase A of az_replay_add has a prefix-sum fast path and an in-order path for windows in which a speculative probe was
    overtaken; both, and odd window sizes, must reproduce the oracle (the parity test re-run in a subprocess with the
    path forced)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_replay_gpu.py"), "-q", "-x", "-k", "test_add_matches_oracle"],
                         env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
