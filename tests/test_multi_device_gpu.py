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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_generation_loop_two_ranks():
    """BASELINE config 5 over NCCL: games sharded over two ranks (exactly 128 complete games each), samples all-gathered device
    to device into both replicas of the replay buffer, data-parallel training with a gradient all-reduce
    (tools/run_generations.py under torch.distributed.run).  The replicas (weights + replay entries) must stay identical."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(root, "tools", "run_generations.py"),
           "--games", "128", "--sims", "32", "--iterations", "2", "--min-replay", "1000", "--eval-games", "0", "--all-ranks"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    # both ranks write to one pipe: parse every JSON object, wherever the line breaks fell
    lines, dec, text, pos = [], json.JSONDecoder(), out.stdout, 0
    while True:
        pos = text.find('{"iteration"', pos)
        if pos < 0:
            break
        obj, end = dec.raw_decode(text, pos)
        lines.append(obj)
        pos = end
    assert len(lines) == 4 and all(m["n_ranks"] == 2 and m["trained"] and m["games"] == 256 for m in lines)
    for it in (0, 1):
        a, b = [m for m in lines if m["iteration"] == it]
        assert {a["rank"], b["rank"]} == {0, 1}
        assert a["replica_digest"] == b["replica_digest"], (it, a, b)
        assert a["positions"] == b["positions"] > 2 * 128 * 20 and a["replay_buffer_size"] == b["replay_buffer_size"]
        assert a["sample_gather"].startswith("device")
    first = [m for m in lines if m["iteration"] == 0][0]
    assert first["replay_buffer_size"] == first["new_unique_states"]
