#!/usr/bin/env python
"""Short driver for ncu captures of the chess-rule kernels and the current search kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402,F401
import alphazero_chess_b200 as az  # noqa: E402
from helpers import KIWIPETE, random_playouts  # noqa: E402

eng = az.Engine(max_games=4096, max_batch=65536, num_simulations=32)
positions, _ = random_playouts(2048, seed=42, max_plies=80)
batch = np.tile(positions, 32)[:65536]
for _ in range(2):
    eng.movegen(batch)
    eng.encode(batch[:16384])
print(int(eng.perft(az.position_from_fen(KIWIPETE), 5)[0]))
eng.set_evaluator_stub(1, 1)
roots = np.repeat(np.array([az.start_position()], az.POSITION_DTYPE), 4096)
eng.search(roots, num_simulations=24, noise_game_ids=np.arange(4096, dtype=np.uint64))
eng.close()
