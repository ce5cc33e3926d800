"""run_all_episodes semantics on the engine (training.rs:340-378): exactly N games played to completion, nothing dropped
when the sample queue is full (finished games park), TEMPERATURE as a parameter (tree.rs:173-177), the evaluation cache
with eviction (parameters.rs:4, moka capacity) and the device-to-device sample path into the replay buffer."""
import os
import subprocess
import sys

import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import orc

pytestmark = pytest.mark.gpu


def _play_generation(e, slots, first_id, total, chunk=64, drain=True, max_calls=4000):
    e.selfplay_begin(slots, first_game_id=first_id, total_games=total)
    out, st = [], None
    for _ in range(max_calls):
        st = e.selfplay_step(chunk)
        if drain and st.pending_samples:
            out.append(e.selfplay_drain())
        if st.active_games == 0 and (not drain or st.pending_samples == 0):
            break
    assert st.active_games == 0
    if not drain:
        out.append(e.selfplay_drain())
    return (np.concatenate(out) if out else np.zeros(0, az.SAMPLE_DTYPE)), st


def _check_games_vs_oracle(samples, game_ids, sims, seed, stub_seed, temperature=1.0):
    prm = orc.make_params(num_simulations=sims, seed=seed, temperature=temperature)
    ev = orc.make_evaluator("stub", stub_seed=stub_seed)
    for gid in game_ids:
        rec = samples[samples["game_id"] == gid]
        rec = rec[np.argsort(rec["ply"])]
        ep = orc.selfplay_episode(prm, ev, game_id=int(gid), max_steps=512)
        assert len(rec) == ep["stats"].n_steps, gid
        assert list(rec["ply"]) == list(range(len(rec))), gid
        for k in range(len(rec)):
            assert rec[k]["position"].tobytes() == ep["positions"][k].tobytes(), (gid, k)
            dense = np.zeros(4096, np.float32)
            dense[rec[k]["index"][: rec[k]["n_visits"]]] = rec[k]["count"][: rec[k]["n_visits"]]
            assert np.array_equal(dense, ep["visits"][k]), (gid, k)
            assert rec[k]["action"] == ep["action"][k], (gid, k)
            assert rec[k]["final_value"] == ep["final_value"][k], (gid, k)


def test_exactly_n_games_are_played_to_completion():
    """20 games on 8 slots: the drained records are the COMPLETE games first_id .. first_id + 19 and nothing else
    (no restart beyond the budget, no game cut off), each identical to the oracle's run_episode."""
    sims, seed, stub_seed, first = 16, 42, 5, 700
    with az.Engine(max_games=8, num_simulations=sims, seed=seed) as e:
        e.set_evaluator_stub(1, stub_seed)
        samples, st = _play_generation(e, 8, first, 20)
        assert st.games_finished == 20 and st.parked_games == 0
        assert st.positions == len(samples)
        assert st.sum_search_depth == int(samples["search_depth"].sum())   # avg_search_depth of training.rs:91-97
        # a further step changes nothing: every slot is idle
        st2 = e.selfplay_step(32)
        assert st2.simulations == st.simulations and st2.active_games == 0
    assert sorted(set(int(g) for g in samples["game_id"])) == list(range(first, first + 20))
    _check_games_vs_oracle(samples, range(first, first + 20), sims, seed, stub_seed)


def test_fewer_games_than_slots():
    with az.Engine(max_games=16, num_simulations=8) as e:
        e.set_evaluator_stub(1, 3)
        samples, st = _play_generation(e, 16, 0, 5)
        assert st.games_finished == 5
        assert sorted(set(int(g) for g in samples["game_id"])) == [0, 1, 2, 3, 4]


_PARK_SCRIPT = """
import sys
sys.path.insert(0, {tests!r})
import _pkg  # noqa: F401
import numpy as np
import alphazero_chess_b200 as az
with az.Engine(max_games=24, num_simulations=8, seed=9) as e:
    e.set_evaluator_stub(1, 4)
    e.selfplay_begin(24, first_game_id=50, total_games=40)
    out, parked_seen, peak = [], 0, 0
    for call in range(6000):
        st = e.selfplay_step(32)
        parked_seen = max(parked_seen, int(st.parked_games))
        peak = max(peak, int(st.pending_samples))
        if call % 40 == 39 or (st.active_games == st.parked_games and st.pending_samples):   # drain rarely: the queue must fill up
            out.append(e.selfplay_drain())
        if st.active_games == 0 and st.pending_samples == 0:
            break
    s = np.concatenate(out)
    np.save({npy!r}, s.view(np.uint8))
    print("RESULT", parked_seen, peak, int(st.games_finished), len(s))
"""


def test_full_sample_queue_parks_games_and_loses_nothing(tmp_path):
    """With a sample queue of 600 records (AZ_SAMPLE_CAP) and a host that rarely drains, finished games park until there is
    room; the queue counter never exceeds the capacity and all 40 games arrive complete and identical to the oracle."""
    here = os.path.dirname(os.path.abspath(__file__))
    npy = str(tmp_path / "parked.npy")
    env = dict(os.environ, AZ_SAMPLE_CAP="600")
    out = subprocess.run([sys.executable, "-c", _PARK_SCRIPT.format(tests=here, npy=npy)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    parked_seen, peak, finished, n = [int(x) for x in [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()[1:]]
    assert parked_seen > 0, "the test did not exercise parking"
    assert peak <= 600 and finished == 40
    samples = np.load(npy).view(az.SAMPLE_DTYPE)
    assert len(samples) == n
    assert sorted(set(int(g) for g in samples["game_id"])) == list(range(50, 90))
    _check_games_vs_oracle(samples, range(50, 90, 3), 8, 9, 4)


@pytest.mark.parametrize("temperature", [0.5, 2.0])
def test_temperature_parameter(temperature):
    """TEMPERATURE != 1 (parameters.rs:33, tree.rs:173-177): move selection uses visits^(1/T) / sum; complete games equal the
    oracle's, and the replay buffer stores the improved policy visits^(1/T) / sum of each EpisodeStep."""
    sims, seed, stub_seed = 20, 13, 21
    with az.Engine(max_games=8, num_simulations=sims, seed=seed, temperature=temperature) as e:
        e.set_evaluator_stub(1, stub_seed)
        samples, st = _play_generation(e, 8, 0, 8)
        assert st.games_finished == 8
        _check_games_vs_oracle(samples, range(8), sims, seed, stub_seed, temperature=temperature)
        rp = az.ReplayBuffer(e, capacity=4096, max_batch=64)
        one = samples[samples["game_id"] == 3]
        one = one[np.argsort(one["ply"])][:24]
        rp.add(one)
        for rec in one[[0, 5, len(one) - 1]]:
            dense = np.zeros(4096, np.float32)
            dense[rec["index"][: rec["n_visits"]]] = rec["count"][: rec["n_visits"]]
            pol, val, cnt = rp.get(rec["position"])
            assert cnt == 1 and val == rec["final_value"]
            assert np.array_equal(pol, orc.improved_policy(dense, temperature))
        rp.close()


def test_replay_policy_normalised_by_the_samples_own_visit_sum():
    """ReplayBuffer::add receives improved_policy = visits / sum(visits) (tree.rs:173-175): samples of a search with another
    simulation count than az_config.num_simulations still give rows that sum to 1."""
    with az.Engine(max_games=4, num_simulations=64) as e:
        s = np.zeros(1, az.SAMPLE_DTYPE)
        s["position"] = az.start_position()
        s["n_visits"] = 3
        s["index"][0, :3] = [100, 588, 1540]
        s["count"][0, :3] = [5, 20, 15]   # a 40-simulation search
        s["final_value"] = 0.25
        rp = az.ReplayBuffer(e, capacity=64, max_batch=16)
        assert rp.add(s) == 1
        pol, val, cnt = rp.get(s["position"][0])
        want = np.zeros(4096, np.float32)
        want[[100, 588, 1540]] = np.array([5, 20, 15], np.float32) / np.float32(40)
        assert np.array_equal(pol, want) and val == 0.25 and cnt == 1
        rp.close()


def test_cache_eviction_is_transparent():
    """A 1024-slot evaluation cache under 16 games x 24 simulations must evict (capacity management like moka's,
    parameters.rs:4); records stay identical to the oracle: a hit returns exactly what the evaluation would."""
    sims, seed, stub_seed = 24, 42, 17
    with az.Engine(max_games=16, num_simulations=sims, seed=seed, cache_log2=10) as e:
        e.set_evaluator_stub(1, stub_seed)
        samples, st = _play_generation(e, 16, 0, 24)
        assert st.games_finished == 24
        assert st.cache_hits > 0 and st.cache_evictions > 0, (st.cache_hits, st.cache_evictions)
        assert st.cache_hits + st.evaluations + st.terminal_leaves >= st.simulations
    _check_games_vs_oracle(samples, range(0, 24, 2), sims, seed, stub_seed)


def test_device_to_device_sample_path():
    """az_selfplay_drain_dev + az_replay_add_dev (the legs either side of the NCCL gather) give the same replay buffer as the
    host path az_selfplay_drain + az_replay_add."""
    import torch

    sims = 12

    def run(dev_path):
        with az.Engine(max_games=16, num_simulations=sims, seed=5) as e:
            e.set_evaluator_stub(1, 8)
            rp = az.ReplayBuffer(e, capacity=3000, max_batch=128)
            e.selfplay_begin(16, first_game_id=0, total_games=24)
            buf = torch.empty(4096 * az.SAMPLE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            total = uniq = 0
            while True:
                st = e.selfplay_step(48)
                if st.pending_samples:
                    if dev_path:
                        n = e.selfplay_drain_dev(buf.data_ptr(), 4096)
                        uniq += rp.add_dev(buf.data_ptr(), n)
                    else:
                        s = e.selfplay_drain()
                        n = len(s)
                        uniq += rp.add(s)
                    total += n
                elif st.active_games == 0:
                    break
            pos, pol, val, vis = rp.export(0, 128)
            n_entries = len(rp)
            rp.close()
            return total, uniq, n_entries, pos.tobytes(), pol.tobytes(), val.tobytes(), vis.tobytes()

    a, b = run(False), run(True)
    assert a[:3] == b[:3] and a[0] > 0
    assert a[3:] == b[3:]


def test_search_history_must_end_in_the_root():
    with az.Engine(max_games=4, num_simulations=8) as e:
        e.set_evaluator_stub(1, 1)
        start = orc.startpos()
        other = orc.play_encoded(start, 12 | (28 << 6))   # e2e4
        hist = np.array([start, other], orc.POSITION_DTYPE)
        e.search(other, num_simulations=8, history=hist, hist_offsets=np.array([0, 2], np.uint32))
        with pytest.raises(az.EngineError):
            e.search(start, num_simulations=8, history=hist, hist_offsets=np.array([0, 2], np.uint32))
