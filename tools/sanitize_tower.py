#!/usr/bin/env python
"""One 4096-board network forward (input convolution, fused 20-layer tower with lazy publication and two board ranges,
heads) for `compute-sanitizer --tool racecheck|synccheck|memcheck`.  No torch import: the engine only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402,F401
import alphazero_chess_b200 as az  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pos = np.repeat(np.array([az.start_position()], az.POSITION_DTYPE), n)
pos["halfmoves"] = np.arange(n) % 90
with az.Engine(max_games=n, precision=0) as e:
    e.load_weights(az.random_weights(seed=3, randomize_bn=True))
    pol, val = e.forward(pos)
    print("forward ok", n, float(pol.sum()), float(np.abs(val).max()))
