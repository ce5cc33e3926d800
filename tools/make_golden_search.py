#!/usr/bin/env python
"""Generates tests/golden/search_stub.json with the ORACLE (CPU): visit counts of tree.rs searches under the shared synthetic
evaluator, with and without Dirichlet noise, and the digest of one complete self-play game.  The fixture freezes today's
agreed behaviour so that a later change cannot move the oracle and the CUDA engine together unnoticed.

    python tools/make_golden_search.py            # rewrites the fixture
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import orc  # noqa: E402

FENS = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
    "rnbqkbnr/ppp1pppp/8/8/3pP3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 3",
]
SEED, SIMS = 5, 64


def main():
    ev = orc.make_evaluator("stub", stub_seed=SEED)
    prm = orc.make_params(num_simulations=SIMS)
    searches = []
    for k, fen in enumerate(FENS):
        for noise in (False, True):
            v, s, d, _ = orc.search(orc.from_fen(fen), prm, ev, noise_game=(100 + k) if noise else -1, noise_ply=k if noise else 0)
            nz = np.flatnonzero(v)
            searches.append({"fen": fen, "noise_game": (100 + k) if noise else -1, "noise_ply": k if noise else 0, "depth": int(d),
                             "visits": [[int(i), int(v[i])] for i in nz],
                             "scores_sha256": hashlib.sha256(s.tobytes()).hexdigest()})
    ep = orc.selfplay_episode(orc.make_params(num_simulations=32, seed=42), ev, game_id=3, max_steps=512)
    n = ep["stats"].n_steps
    game = {"game_id": 3, "sims": 32, "seed": 42, "steps": int(n), "actions": [int(a) for a in ep["action"][:n]],
            "final_value_first": float(ep["final_value"][0]),
            "visits_sha256": hashlib.sha256(ep["visits"][:n].tobytes()).hexdigest(),
            "positions_sha256": hashlib.sha256(ep["positions"][:n].tobytes()).hexdigest()}
    out = {"_how": "python tools/make_golden_search.py (oracle, stub evaluator seed %d)" % SEED, "stub_seed": SEED, "sims": SIMS,
           "searches": searches, "game": game}
    with open(os.path.join(ROOT, "tests", "golden", "search_stub.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(searches), "searches and a game of", n, "plies")


if __name__ == "__main__":
    main()
