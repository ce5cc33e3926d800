"""training.rs run_episode on the GPU against the oracle: with the shared synthetic evaluator and the shared counter-based
generator the complete game records (positions, visit counts, moves, depths, back-filled values) are identical."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import orc

pytestmark = pytest.mark.gpu


def _collect(e, n_games, target_games, max_waves=40000, chunk=64):
    e.selfplay_begin(n_games, first_game_id=0)
    samples = []
    waves = 0
    while waves < max_waves:
        st = e.selfplay_step(chunk)
        waves += chunk
        if st.pending_samples:
            samples.append(e.selfplay_drain())
        if st.games_finished >= target_games:
            break
    return np.concatenate(samples) if samples else np.zeros(0, az.SAMPLE_DTYPE), st


@pytest.mark.parametrize("cache_log2", [0, 14])
def test_selfplay_records_identical_to_oracle(cache_log2):
    sims, seed, stub_seed = 24, 42, 17
    with az.Engine(max_games=16, num_simulations=sims, seed=seed, cache_log2=cache_log2) as e:
        e.set_evaluator_stub(1, stub_seed)
        samples, st = _collect(e, 16, 16)
        assert st.games_finished >= 16 and st.positions >= len(samples)
        # the evaluation cache (training.rs:342) is semantically transparent: same records, fewer evaluations
        assert (st.cache_hits > 0) == (cache_log2 > 0)
        assert st.cache_hits + st.evaluations + st.terminal_leaves >= st.simulations
    prm = orc.make_params(num_simulations=sims, seed=seed)
    ev = orc.make_evaluator("stub", stub_seed=stub_seed)
    finished = sorted(set(int(g) for g in samples["game_id"]))
    checked = 0
    for gid in finished[:12]:
        rec = samples[samples["game_id"] == gid]
        rec = rec[np.argsort(rec["ply"])]
        ep = orc.selfplay_episode(prm, ev, game_id=gid, max_steps=512)
        assert len(rec) == ep["stats"].n_steps, gid
        for k in range(len(rec)):
            assert rec[k]["position"].tobytes() == ep["positions"][k].tobytes(), (gid, k)
            assert np.array_equal(az.improved_policy(rec[k], sims) * np.float32(sims), ep["visits"][k]), (gid, k)
            assert rec[k]["action"] == ep["action"][k], (gid, k)
            assert rec[k]["search_depth"] == ep["depth"][k], (gid, k)
            assert rec[k]["final_value"] == ep["final_value"][k], (gid, k)
        checked += 1
    assert checked >= 8


def test_selfplay_counters_and_network_path():
    w = az.random_weights(seed=5)
    with az.Engine(max_games=64, num_simulations=16) as e:
        e.load_weights(w)
        e.selfplay_begin(64)
        st = e.selfplay_step(40)
        assert st.simulations >= 64 * 30 and st.evaluations > 0
        assert st.positions >= 64  # every game has made at least one move after 40 waves of 16-simulation searches
        s = e.selfplay_drain()
        assert len(s) == st.pending_samples


def test_sharded_games_equal_unsharded():
    """Multi-GPU parity on one device: a game's record depends only on (seed, game id), so two engines that own disjoint
    id ranges (what two ranks do) produce exactly the records one engine produces for the union."""
    sims, stub_seed = 16, 23

    def run(first_id, n):
        with az.Engine(max_games=n, num_simulations=sims, seed=7) as e:
            e.set_evaluator_stub(1, stub_seed)
            e.selfplay_begin(n, first_game_id=first_id)
            out = []
            for _ in range(400):
                st = e.selfplay_step(64)
                if st.pending_samples:
                    out.append(e.selfplay_drain())
                if st.games_finished >= 2 * n:
                    break
            return np.concatenate(out)

    whole = run(100, 8)
    parts = np.concatenate([run(100, 4), run(104, 4)])

    def records(samples, gid):
        r = samples[samples["game_id"] == gid]
        r = r[np.argsort(r["ply"], kind="stable")]
        # a finished game restarts under the next free id, so two shards may both have played the same id: the copies
        # must be identical (the record depends on the id only); keep one
        keep = [0] if len(r) else []
        for k in range(1, len(r)):
            if r[k]["ply"] == r[k - 1]["ply"]:
                assert r[k].tobytes() == r[k - 1].tobytes()
            else:
                keep.append(k)
        return r[keep]

    compared = 0
    for gid in range(100, 108):
        a, b = records(whole, gid), records(parts, gid)
        if len(a) == 0 or len(b) == 0:
            continue
        assert a.tobytes() == b.tobytes(), gid
        compared += 1
    assert compared >= 6


def test_selfplay_with_network_matches_oracle_given_identical_outputs():
    """Whole self-play games through the tcgen05 network: the oracle replays the same games when its tree is fed the
    outputs the GPU network produces for each position (az_forward is batch-invariant, so they are the same numbers)."""
    sims, seed = 12, 99
    w = az.random_weights(seed=11)
    with az.Engine(max_games=4, num_simulations=sims, seed=seed) as e:
        e.load_weights(w)
        e.selfplay_begin(4, first_game_id=500)
        samples = []
        for _ in range(600):
            st = e.selfplay_step(32)
            if st.pending_samples:
                samples.append(e.selfplay_drain())
            if st.games_finished >= 2:
                break
        samples = np.concatenate(samples)

        def cb(ctx, pos_ptr, pol_ptr, val_ptr):
            pos = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * 72).from_address(pos_ptr)).view(az.POSITION_DTYPE)
            p, v = e.forward(pos)
            np.ctypeslib.as_array(pol_ptr, (4096,))[:] = p[0]
            val_ptr[0] = float(v[0])

        ev = orc.make_evaluator("callback", callback=orc.EVAL_FN(cb))
        prm = orc.make_params(num_simulations=sims, seed=seed)
        checked = 0
        for gid in sorted(set(int(g) for g in samples["game_id"]))[:2]:
            rec = samples[samples["game_id"] == gid]
            rec = rec[np.argsort(rec["ply"])]
            ep = orc.selfplay_episode(prm, ev, game_id=gid, max_steps=512)
            assert len(rec) == ep["stats"].n_steps, gid
            for k in range(len(rec)):
                assert rec[k]["position"].tobytes() == ep["positions"][k].tobytes(), (gid, k)
                assert np.array_equal(az.improved_policy(rec[k], sims) * np.float32(sims), ep["visits"][k]), (gid, k)
                assert rec[k]["action"] == ep["action"][k] and rec[k]["final_value"] == ep["final_value"][k], (gid, k)
            checked += 1
        assert checked == 2
