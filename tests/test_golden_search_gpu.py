"""The CUDA engine against the frozen search fixture (tests/golden/search_stub.json): visit counts, accumulated scores,
depths and one complete self-play game, independent of the live oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import ROOT

pytestmark = pytest.mark.gpu

with open(os.path.join(ROOT, "tests", "golden", "search_stub.json")) as f:
    GOLD = json.load(f)


def test_engine_reproduces_frozen_searches():
    with az.Engine(max_games=32, num_simulations=GOLD["sims"]) as e:
        e.set_evaluator_stub(1, GOLD["stub_seed"])
        for noise in (False, True):
            gs = [g for g in GOLD["searches"] if (g["noise_game"] >= 0) == noise]
            roots = np.array([az.position_from_fen(g["fen"]) for g in gs], az.POSITION_DTYPE)
            kw = {}
            if noise:
                kw = {"noise_game_ids": np.array([g["noise_game"] for g in gs], np.uint64),
                      "noise_plies": np.array([g["noise_ply"] for g in gs], np.uint32)}
            visits, scores, depth = e.search(roots, num_simulations=GOLD["sims"], want_scores=True, **kw)
            for k, g in enumerate(gs):
                want = np.zeros(4096, np.float32)
                for i, c in g["visits"]:
                    want[i] = c
                assert np.array_equal(visits[k], want), (g["fen"], noise)
                assert depth[k] == g["depth"]
                assert hashlib.sha256(np.ascontiguousarray(scores[k]).tobytes()).hexdigest() == g["scores_sha256"]


def test_engine_reproduces_frozen_game():
    g = GOLD["game"]
    with az.Engine(max_games=4, num_simulations=g["sims"], seed=g["seed"]) as e:
        e.set_evaluator_stub(1, GOLD["stub_seed"])
        e.selfplay_begin(4, first_game_id=0)
        chunks = []
        for _ in range(2000):
            st = e.selfplay_step(32)
            if st.pending_samples:
                chunks.append(e.selfplay_drain())
                if any((c["game_id"] == g["game_id"]).any() for c in chunks):
                    break
        rec = np.concatenate(chunks)
        rec = rec[rec["game_id"] == g["game_id"]]
        rec = rec[np.argsort(rec["ply"])]
    assert len(rec) == g["steps"] and [int(a) for a in rec["action"]] == g["actions"]
    visits = np.stack([az.improved_policy(r, g["sims"]) * np.float32(g["sims"]) for r in rec]).astype(np.float32)
    assert hashlib.sha256(visits.tobytes()).hexdigest() == g["visits_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(rec["position"]).tobytes()).hexdigest() == g["positions_sha256"]
    assert float(rec["final_value"][0]) == g["final_value_first"]


def test_engine_reproduces_frozen_move_lists():
    with open(os.path.join(ROOT, "tests", "golden", "movegen.json")) as f:
        gold = json.load(f)["positions"]
    pos = np.array([az.position_from_fen(g["fen"]) for g in gold], az.POSITION_DTYPE)
    with az.Engine(max_games=64) as e:
        moves, index, count = e.movegen(pos)
        planes = e.encode(pos)
    for k, g in enumerate(gold):
        assert count[k] == len(g["moves"]), g["fen"]
        assert [int(m) for m in moves[k, : count[k]]] == g["moves"], g["fen"]
        assert [int(i) for i in index[k, : count[k]]] == g["index"], g["fen"]
        assert hashlib.sha256(np.ascontiguousarray(planes[k]).tobytes()).hexdigest() == g["planes_sha256"], g["fen"]
