"""Two engines on two GPUs inside one process (kernel attributes and tensor maps are per device)."""
import numpy as np
import pytest
import torch

import alphazero_chess_b200 as az
from helpers import orc, random_playouts

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_one_process():
    w = az.random_weights(seed=2)
    positions, _ = random_playouts(64, seed=6, max_plies=60)
    with az.Engine(device=0, max_games=64, num_simulations=16) as e0, az.Engine(device=1, max_games=64, num_simulations=16) as e1:
        e0.load_weights(w)
        e1.load_weights(w)
        p0, v0 = e0.forward(positions)
        p1, v1 = e1.forward(positions)
        assert np.array_equal(p0, p1) and np.array_equal(v0, v1)
        roots = positions[[i for i in range(64) if orc.outcome(positions[i]) == 0][:16]]
        a, _, _ = e0.search(roots, num_simulations=16)
        b, _, _ = e1.search(roots, num_simulations=16)
        assert np.array_equal(a, b)
        assert int(e1.perft(orc.startpos(), 4)[0]) == 197281
