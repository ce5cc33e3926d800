"""Parity at the size bench.py runs (BASELINE configs[2]: 4096 boards / 4096 games x 800 simulations).

At 4096 boards every CTA pair of conv_tower_kernel owns 7 tiles per range, so the kernel takes the LAZY publication
branch and walks TWO board ranges inside one launch; at <= 256 boards (the other network tests) it owns one tile and
publishes eagerly.  These tests cover that branch, the ragged last tile (4095, 2073 boards) and the search kernels at
4096 roots x 800 simulations.  Tolerances: 1e-2 absolute for the bf16 network (north_star), bit-exact for everything
the search computes."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import orc, random_playouts, torch_reference_forward

pytestmark = pytest.mark.gpu

N_BOARDS = 4096


@pytest.fixture(scope="module")
def weights():
    return az.random_weights(seed=3, randomize_bn=True)


@pytest.fixture(scope="module")
def positions():
    p, _ = random_playouts(N_BOARDS, seed=77, max_plies=120)
    return p


@pytest.fixture(scope="module")
def torch_ref(weights, positions):
    planes = np.stack([orc.to_tensor(p) for p in positions])
    rp, rv = [], []
    for lo in range(0, N_BOARDS, 512):
        a, b = torch_reference_forward(weights, planes[lo:lo + 512])
        rp.append(a)
        rv.append(b)
    return np.concatenate(rp), np.concatenate(rv)


def test_forward_at_bench_size_matches_torch(weights, positions, torch_ref):
    """az_forward at 2072 / 2073 / 4095 / 4096 boards vs plain PyTorch fp32 (agent.rs:112-144), <= 1e-2 absolute, and the
    rows are bit-identical to the same boards evaluated in small batches (eager publication branch)."""
    rp, rv = torch_ref
    with az.Engine(max_games=N_BOARDS, precision=0) as e:
        e.load_weights(weights)
        small_p = np.concatenate([e.forward(positions[lo:lo + 128])[0] for lo in range(0, 512, 128)])
        small_v = np.concatenate([e.forward(positions[lo:lo + 128])[1] for lo in range(0, 512, 128)])
        tail_p, tail_v = e.forward(positions[N_BOARDS - 100:])
        worst = 0.0
        for n in (2072, 2073, 4095, 4096):
            pol, val = e.forward(positions[:n])
            dp, dv = np.abs(pol - rp[:n]).max(), np.abs(val - rv[:n]).max()
            worst = max(worst, dp, dv)
            assert dp <= 1e-2 and dv <= 1e-2, (n, dp, dv)
            assert np.abs(pol - rp[:n]).max() / rp[:n].max() < 0.2, n
            assert np.allclose(pol.sum(1), 1.0, atol=1e-3)
            # batch invariance across the eager (<= 256 boards) and lazy / two-range (>= 2072 boards) branches
            assert np.array_equal(pol[:512], small_p) and np.array_equal(val[:512], small_v), n
            if n == N_BOARDS:
                assert np.array_equal(pol[N_BOARDS - 100:], tail_p) and np.array_equal(val[N_BOARDS - 100:], tail_v)
        print(f"\nbf16 network at 2072..4096 boards: worst |d| vs torch fp32 = {worst:.2e}")


def test_forward_at_bench_size_is_deterministic(weights, positions):
    """compute-sanitizer (racecheck / synccheck) is closed on this GPU pool, so the ordering argument of the lazy publication
    branch is checked the other way round: a race between an epilogue store and the next layer's TMA read would show as
    run-to-run differences.  30 repetitions at 4096 and 4095 boards must give one digest each."""
    with az.Engine(max_games=N_BOARDS, precision=0) as e:
        e.load_weights(weights)
        for n in (4096, 4095):
            seen = set()
            for _ in range(30):
                pol, val = e.forward(positions[:n])
                seen.add(hashlib.sha256(pol.tobytes() + val.tobytes()).hexdigest())
            assert len(seen) == 1, (n, len(seen))


_MODE_SCRIPT = """
import sys, hashlib
sys.path.insert(0, {tests!r})
import _pkg  # noqa: F401
import numpy as np
import alphazero_chess_b200 as az
p = np.load({npy!r}).view(az.POSITION_DTYPE)
with az.Engine(max_games=4096, precision=0) as e:
    e.load_weights(az.random_weights(seed=3, randomize_bn=True))
    for n in (4096, 4095, 2073):
        pol, val = e.forward(p[:n])
        print("DIGEST", n, hashlib.sha256(pol.tobytes() + val.tobytes()).hexdigest())
"""


def test_launch_modes_identical_at_bench_size(positions, tmp_path):
    """AZ_TOWER_FUSED = 0 / 1 / 2 (21 launches, input convolution + fused tower, everything in one launch) and a forced
    three-range walk give the same bits at 4096, 4095 and 2073 boards."""
    here = os.path.dirname(os.path.abspath(__file__))
    digests = []
    npy = str(tmp_path / "positions.npy")
    np.save(npy, positions.view(np.uint8))
    for mode, split in (("0", None), ("1", None), ("2", None), ("1", "3"), ("1", "1")):
        env = dict(os.environ, AZ_TOWER_FUSED=mode)
        if split:
            env["AZ_TOWER_SPLIT"] = split
        out = subprocess.run([sys.executable, "-c", _MODE_SCRIPT.format(tests=here, npy=npy)], env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append(tuple(l for l in out.stdout.splitlines() if l.startswith("DIGEST")))
        assert len(digests[-1]) == 3
    assert len(set(digests)) == 1, digests


def test_search_4096_roots_800_sims_bit_exact_vs_oracle(positions):
    """az_search with the shared synthetic evaluator at the bench's size: 4096 distinct roots x 800 simulations with root
    noise; visits, accumulated scores and depths of 160 sampled roots are bit-equal to the oracle (tree.rs:180-207)."""
    sims, stub_seed = 800, 29
    keep = np.array([i for i in range(N_BOARDS) if orc.outcome(positions[i]) == 0])
    roots = positions[keep]
    n = len(roots)
    ids = np.arange(n, dtype=np.uint64) + 5000
    plies = (np.arange(n) % 11).astype(np.uint32)
    with az.Engine(max_games=N_BOARDS, num_simulations=sims) as e:
        e.set_evaluator_stub(1, stub_seed)
        visits, scores, depth = e.search(roots, num_simulations=sims, noise_game_ids=ids, noise_plies=plies, want_scores=True)
    assert np.all(visits.sum(1) == sims)
    prm = orc.make_params(num_simulations=sims)
    ev = orc.make_evaluator("stub", stub_seed=stub_seed)
    rng = np.random.default_rng(1)
    sample = np.unique(np.concatenate([[0, 1, n - 2, n - 1], rng.choice(n, 160, replace=False)]))
    for i in sample:
        v, s, d, _ = orc.search(roots[i], prm, ev, noise_game=int(ids[i]), noise_ply=int(plies[i]))
        assert np.array_equal(visits[i], v), i
        assert np.array_equal(scores[i], s), i
        assert depth[i] == d, i


def test_selfplay_4096_games_800_sims_records_identical():
    """Self-play at the bench's size with the synthetic evaluator: after 4 plies of 4096 games x 800 simulations the staged
    EpisodeSteps (positions, visit counts, depths, moves) of 48 sampled games equal the oracle's run_episode."""
    sims, stub_seed, seed, plies = 800, 31, 42, 4
    G = N_BOARDS
    with az.Engine(max_games=G, num_simulations=sims, seed=seed) as e:
        e.set_evaluator_stub(1, stub_seed)
        e.selfplay_begin(G, first_game_id=100000)
        st = e.selfplay_step(plies * (sims + 2) + 8)
        assert st.simulations >= G * sims * plies
        rng = np.random.default_rng(2)
        slots = np.unique(np.concatenate([[0, 1, G - 1], rng.choice(G, 48, replace=False)]))
        staged = {int(s): e.selfplay_staged(int(s)) for s in slots}
    prm = orc.make_params(num_simulations=sims, seed=seed)
    ev = orc.make_evaluator("stub", stub_seed=stub_seed)
    for slot, rec in staged.items():
        assert len(rec) >= plies, slot
        gid = int(rec[0]["game_id"])
        assert gid == 100000 + slot
        ep = orc.selfplay_episode(prm, ev, game_id=gid, max_steps=plies)
        for k in range(plies):
            assert rec[k]["position"].tobytes() == ep["positions"][k].tobytes(), (slot, k)
            dense = np.zeros(4096, np.float32)
            dense[rec[k]["index"][: rec[k]["n_visits"]]] = rec[k]["count"][: rec[k]["n_visits"]]
            assert np.array_equal(dense, ep["visits"][k]), (slot, k)
            # improved_policy = visits / S exactly as the reference computes it at T = 1 (tree.rs:173-177)
            assert np.array_equal(az.improved_policy(rec[k], sims), ep["visits"][k] / np.float32(sims)), (slot, k)
            assert rec[k]["action"] == ep["action"][k], (slot, k)
            assert rec[k]["search_depth"] == ep["depth"][k], (slot, k)


def test_network_search_at_bench_size_matches_oracle_given_identical_outputs(weights, positions):
    """The benchmarked path itself (bf16 tcgen05 network, priors scattered by the heads kernel, 4096 roots x 800 simulations):
    the oracle's tree, fed the outputs the GPU network gives for each leaf, reproduces visits and scores bit for bit."""
    sims = 800
    keep = np.array([i for i in range(N_BOARDS) if orc.outcome(positions[i]) == 0])
    roots = positions[keep]
    n = len(roots)
    ids = np.arange(n, dtype=np.uint64) + 9000
    with az.Engine(max_games=N_BOARDS, num_simulations=sims, precision=0) as e:
        e.load_weights(weights)
        visits, scores, depth = e.search(roots, num_simulations=sims, noise_game_ids=ids, want_scores=True)
        assert np.all(visits.sum(1) == sims)

        def cb(ctx, pos_ptr, pol_ptr, val_ptr):
            pos = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * 72).from_address(pos_ptr)).view(az.POSITION_DTYPE)
            p, v = e.forward(pos)
            np.ctypeslib.as_array(pol_ptr, (4096,))[:] = p[0]
            val_ptr[0] = float(v[0])

        ev = orc.make_evaluator("callback", callback=orc.EVAL_FN(cb))
        prm = orc.make_params(num_simulations=sims)
        for i in (0, n // 3, n - 1, 2071, 2072):
            v, s, d, _ = orc.search(roots[i], prm, ev, noise_game=int(ids[i]), noise_ply=0)
            assert np.array_equal(visits[i], v), i
            assert np.array_equal(scores[i], s), i
            assert depth[i] == d, i
    digest = hashlib.sha256(visits.tobytes()).hexdigest()
    print(f"\n4096 x 800 network search: visits digest {digest[:16]}")
