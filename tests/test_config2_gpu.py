"""BASELINE configs[1] at its stated size: 65,536 positions (seeded random playouts from the start position, Kiwipete and
the special positions) through every chess.rs entry point of the C ABI, bit-exact against the oracle, plus perft depth 6
from both roots against the public tables."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import KIWIPETE, SPECIAL_FENS, orc

pytestmark = pytest.mark.gpu

N = 65536


@pytest.fixture(scope="module")
def eng():
    e = az.Engine(max_games=64, max_batch=N, num_simulations=16)
    yield e
    e.close()


@pytest.fixture(scope="module")
def corpus():
    roots = [orc.from_fen(f) for f in SPECIAL_FENS]
    roots = np.array([orc.startpos(), orc.from_fen(KIWIPETE)] * 8 + [r for r in roots if orc.outcome(r) == 0], orc.POSITION_DTYPE)
    pos, hist, offs = orc.playout_corpus(N, seed=42, max_plies=80, roots=roots)   # seed 42 = parameters.rs:6
    return pos, hist, offs


def test_movegen_65536_positions(eng, corpus):
    pos, _, _ = corpus
    moves, index, count = eng.movegen(pos)
    want_moves, want_index, want_count = orc.legal_moves_batch(pos)
    assert np.array_equal(count, want_count)
    assert np.array_equal(moves, want_moves)                       # ordered lists, 0xFFFF beyond the count
    live = np.arange(256)[None, :] < count[:, None]
    assert np.array_equal(index[live], want_index[live])
    assert count.max() <= 218 and (count == 0).sum() > 0 and int(count.sum()) > 1_500_000


def test_planes_65536_positions(eng, corpus):
    pos, _, _ = corpus
    assert np.array_equal(eng.encode(pos), orc.to_tensor_batch(pos))


def test_play_move_65536_positions(eng, corpus):
    """play_move through a policy index with the full game histories (repetition counting included): 90 % legal indices,
    10 % arbitrary ones (mostly illegal)."""
    pos, hist, offs = corpus
    _, index, count = orc.legal_moves_batch(pos)
    rng = np.random.default_rng(1)
    pick = index[np.arange(N), rng.integers(0, 256, N) % np.maximum(count, 1)]
    act = np.where((rng.random(N) < 0.9) & (count > 0), pick, rng.integers(0, 4096, N)).astype(np.uint16)
    got_pos, got_res = eng.play_move(pos, act, hist, offs)
    want_pos, want_res = orc.play_move_batch(pos, act, hist, offs)
    assert np.array_equal(got_res, want_res)
    assert got_pos.tobytes() == want_pos.tobytes()
    seen = set(int(r) for r in np.unique(got_res))
    assert {-1, 0} <= seen and len(seen) >= 3


def test_index_to_move_65536_positions(eng, corpus):
    pos, _, _ = corpus
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 4096, N).astype(np.uint16)
    got = eng.index_to_move(pos, idx)
    want = orc.index_to_move_batch(pos, idx)
    assert np.array_equal(got, want)
    assert (got != az.MOVE_NONE).sum() > 100
    # and move_to_index is its inverse on every legal non-under-promotion move of a sample of positions
    moves, index, count = eng.movegen(pos[:4096])
    for i in range(0, 4096, 37):
        for k in range(count[i]):
            promo = (int(moves[i, k]) >> 12) & 7
            if promo in (0, 4):
                assert eng.index_to_move(pos[i], [index[i, k]])[0] == moves[i, k]


@pytest.mark.parametrize("name,fen,depth,nodes", [
    ("startpos", "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 6, 119_060_324),
    ("kiwipete", KIWIPETE, 5, 193_690_690),
    ("kiwipete", KIWIPETE, 6, 8_031_647_685),
])
def test_perft_depth_5_6(eng, name, fen, depth, nodes):
    assert int(eng.perft(az.position_from_fen(fen), depth)[0]) == nodes


def test_perft_from_65536_roots(eng, corpus):
    """Independent roots share the level buffers: depth 2 from all 65,536 positions against the oracle on the host cores."""
    pos, _, _ = corpus
    got = eng.perft(pos, 2)
    want = orc.perft_batch(pos, 2, 8)
    assert np.array_equal(got, want)
