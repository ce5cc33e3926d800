"""Edge cases of the C ABI: empty and ragged batches, the 218-move position (maximum fan-out), capacity errors."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import orc

pytestmark = pytest.mark.gpu

MAX_MOVES_FEN = "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1"   # 218 legal moves, the known maximum


@pytest.fixture(scope="module")
def eng():
    e = az.Engine(max_games=32, max_batch=256, num_simulations=64)
    e.set_evaluator_stub(1, 3)
    yield e
    e.close()


def test_empty_batches(eng):
    none = np.zeros(0, az.POSITION_DTYPE)
    assert eng.movegen(none)[2].shape == (0,)
    assert eng.perft(none, 3).shape == (0,)
    assert eng.encode(none).shape == (0, 19, 8, 8)
    assert eng.play_move(none, np.zeros(0, np.uint16))[1].shape == (0,)
    assert eng.move_to_index(none, np.zeros(0, np.uint16)).shape == (0,)
    assert eng.index_to_move(none, np.zeros(0, np.uint16)).shape == (0,)
    v, _, d = eng.search(none, num_simulations=8)
    assert v.shape == (0, 4096) and d.shape == (0,)


def test_maximum_fanout_position(eng):
    pos = orc.from_fen(MAX_MOVES_FEN)
    mv, ix = orc.legal_moves(pos)
    assert len(mv) == 218
    moves, index, count = eng.movegen(pos)
    assert count[0] == 218 and np.array_equal(moves[0, :218], mv) and np.array_equal(index[0, :218], ix)
    assert int(eng.perft(pos, 2)[0]) == orc.perft(pos, 2)
    # search with root noise over 218 components and a full-width edge list
    sims = 64
    visits, scores, depth = eng.search(pos, num_simulations=sims, noise_game_ids=np.array([77], np.uint64), noise_plies=np.array([5], np.uint32),
                                       want_scores=True)
    v, s, d, _ = orc.search(pos, orc.make_params(num_simulations=sims), orc.make_evaluator("stub", stub_seed=3), noise_game=77, noise_ply=5)
    assert np.array_equal(visits[0], v) and np.array_equal(scores[0], s) and depth[0] == d


def test_ragged_histories_and_mixed_roots(eng):
    # roots with different history lengths (0, 1, many) in one call
    a = orc.startpos()
    b = orc.from_fen(MAX_MOVES_FEN)
    hist_a = [a]
    p = a
    for _ in range(6):
        mv, _ = orc.legal_moves(p)
        p = orc.play_encoded(p, int(mv[len(mv) // 2]))
        hist_a.append(p)
    roots = np.array([p, b, a], az.POSITION_DTYPE)
    hist = np.array(hist_a + [a], az.POSITION_DTYPE)          # game 0: 7 positions, game 1: none, game 2: itself
    offs = np.array([0, len(hist_a), len(hist_a), len(hist_a) + 1], np.uint32)
    visits, _, _ = eng.search(roots, num_simulations=32, history=hist, hist_offsets=offs)
    prm, ev = orc.make_params(num_simulations=32), orc.make_evaluator("stub", stub_seed=3)
    assert np.array_equal(visits[0], orc.search(p, prm, ev, history=np.array(hist_a, az.POSITION_DTYPE))[0])
    assert np.array_equal(visits[1], orc.search(b, prm, ev)[0])
    assert np.array_equal(visits[2], orc.search(a, prm, ev)[0])


def test_capacity_and_argument_errors(eng):
    with pytest.raises(az.EngineError):
        eng.movegen(np.zeros(257, az.POSITION_DTYPE))          # > max_batch
    with pytest.raises(az.EngineError):
        eng.search(np.repeat(np.array([orc.startpos()], az.POSITION_DTYPE), 33), num_simulations=8)   # > max_games
    with pytest.raises(az.EngineError):
        eng.perft(orc.startpos(), 99)
    with pytest.raises(az.EngineError):
        eng.selfplay_step(1)                                    # before selfplay_begin
    with pytest.raises(az.EngineError):
        az.Engine(max_games=0)
    # an edge pool that is too small is reported, never silently wrong
    with az.Engine(max_games=4, num_simulations=64, edge_capacity_per_node=4) as small:
        small.set_evaluator_stub(1, 3)
        with pytest.raises(az.EngineError):
            small.search(orc.startpos(), num_simulations=64)
