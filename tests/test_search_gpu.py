"""tree.rs on the GPU against the oracle with the shared synthetic evaluator: visit counts and accumulated scores are
bit-exact (one simulation in flight per game, same f32 arithmetic, same move order)."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import SPECIAL_FENS, orc, random_playouts

pytestmark = pytest.mark.gpu

SEED = 11


@pytest.fixture(scope="module")
def eng():
    e = az.Engine(max_games=256, num_simulations=256)
    e.set_evaluator_stub(1, SEED)
    yield e
    e.close()


def _roots(n, seed):
    positions, histories = random_playouts(4 * n, seed=seed, max_plies=100)
    keep = [i for i in range(len(positions)) if orc.outcome(positions[i]) == 0][:n]
    return positions[keep], [histories[i] for i in keep]


def test_stub_evaluator_is_shared(eng):
    # the GPU stub must reproduce the oracle's synthetic policy/value bit for bit: check through a 1-simulation search
    prm = orc.make_params(num_simulations=1)
    ev = orc.make_evaluator("stub", stub_seed=SEED)
    roots, _ = _roots(32, 5)
    visits, _, _ = eng.search(roots, num_simulations=1)
    for i, r in enumerate(roots):
        v, _, _, _ = orc.search(r, prm, ev)
        assert np.array_equal(visits[i], v), i


@pytest.mark.parametrize("sims", [16, 100, 256])
def test_visit_counts_bit_exact(eng, sims):
    roots, histories = _roots(96, 21 + sims)
    hist = np.concatenate(histories)
    offs = np.zeros(len(roots) + 1, np.uint32)
    offs[1:] = np.cumsum([len(h) for h in histories])
    visits, scores, depth = eng.search(roots, num_simulations=sims, history=hist, hist_offsets=offs, want_scores=True)
    prm = orc.make_params(num_simulations=sims)
    ev = orc.make_evaluator("stub", stub_seed=SEED)
    for i, r in enumerate(roots):
        v, s, d, _ = orc.search(r, prm, ev, history=histories[i])
        assert visits[i].sum() == sims
        assert np.array_equal(visits[i], v), i
        assert np.array_equal(scores[i], s), i
        assert depth[i] == d, i


def test_special_positions_and_noise(eng):
    roots = np.array([orc.from_fen(f) for f in SPECIAL_FENS if orc.outcome(orc.from_fen(f)) == 0], orc.POSITION_DTYPE)
    ids = np.arange(len(roots), dtype=np.uint64) + 1000
    plies = np.arange(len(roots), dtype=np.uint32) % 7
    visits, scores, depth = eng.search(roots, num_simulations=64, noise_game_ids=ids, noise_plies=plies, want_scores=True)
    prm = orc.make_params(num_simulations=64)
    ev = orc.make_evaluator("stub", stub_seed=SEED)
    for i, r in enumerate(roots):
        v, s, d, _ = orc.search(r, prm, ev, noise_game=int(ids[i]), noise_ply=int(plies[i]))
        assert np.array_equal(visits[i], v), i
        assert np.array_equal(scores[i], s), i
        assert depth[i] == d


def test_repetition_inside_the_tree(eng):
    # shuffling pieces: the search must see draws by repetition through history + path (chess.rs:52-60, tree.rs:210)
    pos = orc.startpos()
    hist = [pos.copy()]
    for uci in ["g1f3", "g8f6", "f3g1", "f6g8", "g1f3", "g8f6"]:
        f = (ord(uci[0]) - 97) + 8 * (int(uci[1]) - 1)
        t = (ord(uci[2]) - 97) + 8 * (int(uci[3]) - 1)
        pos = orc.play_encoded(pos, f | (t << 6))
        hist.append(pos.copy())
    h = np.array(hist, orc.POSITION_DTYPE)
    visits, scores, depth = eng.search(pos, num_simulations=200, history=h, hist_offsets=np.array([0, len(h)], np.uint32), want_scores=True)
    v, s, d, _ = orc.search(pos, orc.make_params(num_simulations=200), orc.make_evaluator("stub", stub_seed=SEED), history=h)
    assert np.array_equal(visits[0], v) and np.array_equal(scores[0], s) and depth[0] == d


def test_search_invariants_and_errors(eng):
    roots, _ = _roots(8, 99)
    visits, _, depth = eng.search(roots, num_simulations=50)
    assert np.all(visits.sum(axis=1) == 50) and np.all(depth >= 1)
    with pytest.raises(az.EngineError):
        eng.search(roots, num_simulations=100000)
    mate = orc.from_fen("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1")
    v, _, _ = eng.search(mate, num_simulations=10)
    assert v.sum() == 0
