#!/usr/bin/env python
"""BASELINE config 2: batched movegen and perft on the GPU next to the oracle on the host cores.
Prints one JSON line per measurement (wall clock around the C-ABI call, host buffers in and out)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402,F401
import alphazero_chess_b200 as az  # noqa: E402
from helpers import KIWIPETE, orc, random_playouts  # noqa: E402

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
KAT = {("startpos", 5): 4865609, ("startpos", 6): 119060324, ("kiwipete", 4): 4085603, ("kiwipete", 5): 193690690,
       ("kiwipete", 6): 8031647685}


def main():
    cores = os.cpu_count() or 1
    eng = az.Engine(max_games=64, max_batch=65536, num_simulations=16)
    positions, _ = random_playouts(4096, seed=42, max_plies=80)
    batch = np.tile(positions, 16)[:65536]
    eng.movegen(batch[:1024])
    for _ in range(2):
        t0 = time.perf_counter()
        moves, index, count = eng.movegen(batch)
        dt = time.perf_counter() - t0
    print(json.dumps({"bench": "movegen", "positions": int(len(batch)), "moves": int(count.sum()), "seconds": dt,
                      "positions_per_sec": len(batch) / dt, "note": "az_movegen incl. H2D of 72 B/position and D2H of 1028 B/position"}))
    t0 = time.perf_counter()
    n_cpu = 8192
    for p in batch[:n_cpu]:
        orc.legal_moves(p)
    dtc = time.perf_counter() - t0
    print(json.dumps({"bench": "movegen_cpu_oracle", "positions": n_cpu, "seconds": dtc, "positions_per_sec": n_cpu / dtc, "cores": 1,
                      "note": "python loop over the oracle's legal_moves (ctypes overhead included)"}))
    for name, fen in (("startpos", START), ("kiwipete", KIWIPETE)):
        pos = az.position_from_fen(fen)
        for depth in (5, 6):
            want = KAT.get((name, depth))
            if want is None or (name == "kiwipete" and depth == 6 and "--deep" not in sys.argv):
                continue
            eng.perft(pos, 3)
            t0 = time.perf_counter()
            got = int(eng.perft(pos, depth)[0])
            dt = time.perf_counter() - t0
            assert got == want, (name, depth, got, want)
            print(json.dumps({"bench": "perft", "position": name, "depth": depth, "nodes": got, "seconds": dt, "nodes_per_sec": got / dt}))
    # a batch of roots: perft 3 from 65,536 positions (independent roots share the level buffers)
    t0 = time.perf_counter()
    nodes = eng.perft(batch, 3)
    dt = time.perf_counter() - t0
    print(json.dumps({"bench": "perft_batch", "roots": int(len(batch)), "depth": 3, "nodes": int(nodes.sum()), "seconds": dt,
                      "nodes_per_sec": float(nodes.sum()) / dt}))
    # oracle on all host cores: perft 5 from the two roots split over depth-1 children
    for name, fen, depth in (("startpos", START, 5), ("kiwipete", KIWIPETE, 4)):
        root = orc.from_fen(fen)
        mv, _ = orc.legal_moves(root)
        kids = np.array([orc.play_encoded(root, int(m)) for m in mv], orc.POSITION_DTYPE)
        t0 = time.perf_counter()
        total = int(orc.perft_batch(kids, depth - 1, cores).sum())
        dt = time.perf_counter() - t0
        assert total == KAT[(name, depth)]
        print(json.dumps({"bench": "perft_cpu_oracle", "position": name, "depth": depth, "nodes": total, "seconds": dt,
                          "nodes_per_sec": total / dt, "cores": cores}))
    # Player::MiniMax(4) (chess.rs:247-318): one move decision for each of 256 positions, full width, four plies
    mm_pos = positions[:256]
    eng.minimax(mm_pos[:8], 2)
    for depth in (3, 4):
        t0 = time.perf_counter()
        scores, cnt = eng.minimax(mm_pos, depth)
        dt = time.perf_counter() - t0
        leaves = int(eng.perft(mm_pos, depth).sum())  # upper bound: finished games are not expanded by negamax
        print(json.dumps({"bench": "minimax", "roots": int(len(mm_pos)), "depth": depth, "seconds": dt, "decisions_per_sec": len(mm_pos) / dt,
                          "perft_leaves": leaves, "leaves_per_sec": leaves / dt}))
    t0 = time.perf_counter()
    n_mm_cpu = 4
    for p in mm_pos[:n_mm_cpu]:
        orc.minimax_scores(p, 4)
    dt = time.perf_counter() - t0
    print(json.dumps({"bench": "minimax_cpu_oracle", "roots": n_mm_cpu, "depth": 4, "seconds": dt, "decisions_per_sec": n_mm_cpu / dt, "cores": 1}))
    eng.close()
    # replay buffer (memory.rs): order-exact add and batch sampling
    e2 = az.Engine(max_games=512, num_simulations=16, seed=3)
    e2.set_evaluator_stub(1, 8)
    e2.selfplay_begin(512)
    chunks = []
    for _ in range(400):
        st = e2.selfplay_step(64)
        if st.pending_samples:
            chunks.append(e2.selfplay_drain())
        if st.games_finished >= 512:
            break
    samples = np.concatenate(chunks)
    rb = az.ReplayBuffer(e2, capacity=100_000, max_batch=512)
    t0 = time.perf_counter()
    nu = rb.add(samples)
    dt = time.perf_counter() - t0
    print(json.dumps({"bench": "replay_add", "steps": int(len(samples)), "new_unique": int(nu), "seconds": dt, "steps_per_sec": len(samples) / dt,
                      "note": "az_replay_add: H2D of 1120 B/step + in-order resolve by one warp (32 probes at a time) + one CTA per touched slot (bit-exact running means)"}))
    rb.sample(512, seed=0)
    t0 = time.perf_counter()
    for k in range(50):
        rb.sample(512, seed=k)
    dt = (time.perf_counter() - t0) / 50
    print(json.dumps({"bench": "replay_sample", "batch": 512, "seconds": dt, "batches_per_sec": 1 / dt,
                      "note": "az_replay_sample incl. D2H of planes + dense policy rows (10.9 MB per batch)"}))
    rb.close()
    e2.close()


if __name__ == "__main__":
    main()
