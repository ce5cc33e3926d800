"""The oracle against the frozen search fixture (tests/golden/search_stub.json, made by tools/make_golden_search.py)."""
import hashlib
import json
import os

import numpy as np

from helpers import ROOT, orc

with open(os.path.join(ROOT, "tests", "golden", "search_stub.json")) as f:
    GOLD = json.load(f)


def test_oracle_reproduces_frozen_searches():
    ev = orc.make_evaluator("stub", stub_seed=GOLD["stub_seed"])
    prm = orc.make_params(num_simulations=GOLD["sims"])
    for g in GOLD["searches"]:
        v, s, d, _ = orc.search(orc.from_fen(g["fen"]), prm, ev, noise_game=g["noise_game"], noise_ply=g["noise_ply"])
        want = np.zeros(4096, np.float32)
        for i, c in g["visits"]:
            want[i] = c
        assert np.array_equal(v, want), g["fen"]
        assert d == g["depth"] and hashlib.sha256(s.tobytes()).hexdigest() == g["scores_sha256"]
        assert v.sum() == GOLD["sims"]


def test_oracle_reproduces_frozen_game():
    g = GOLD["game"]
    ev = orc.make_evaluator("stub", stub_seed=GOLD["stub_seed"])
    ep = orc.selfplay_episode(orc.make_params(num_simulations=g["sims"], seed=g["seed"]), ev, game_id=g["game_id"], max_steps=512)
    n = ep["stats"].n_steps
    assert n == g["steps"] and [int(a) for a in ep["action"][:n]] == g["actions"]
    assert hashlib.sha256(ep["visits"][:n].tobytes()).hexdigest() == g["visits_sha256"]
    assert hashlib.sha256(ep["positions"][:n].tobytes()).hexdigest() == g["positions_sha256"]
    assert float(ep["final_value"][0]) == g["final_value_first"]


def test_oracle_reproduces_frozen_move_lists():
    with open(os.path.join(ROOT, "tests", "golden", "movegen.json")) as f:
        gold = json.load(f)["positions"]
    assert len(gold) >= 20
    for g in gold:
        p = orc.from_fen(g["fen"])
        mv, idx = orc.legal_moves(p)
        assert [int(m) for m in mv] == g["moves"] and [int(i) for i in idx] == g["index"], g["fen"]
        assert orc.outcome(p) == g["outcome"]
        assert hashlib.sha256(orc.to_tensor(p).tobytes()).hexdigest() == g["planes_sha256"]
