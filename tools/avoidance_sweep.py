#!/usr/bin/env python
"""How far does evaluation avoidance go?  (a measured route to the 1e8 simulations/s target.)

Trains the 10x128 network for a few short generations on this GPU (the repo's own loop), and after selected generations
measures BASELINE configs[2] self-play (4096 games x 800 sims) with the evaluation cache on: simulations/s, evaluations/s and
the avoided fraction.  A second sweep sharpens the trained policy head (logits x k): a synthetic stand-in for a strong
network's peaked priors.  One JSON line per measurement.

    python tools/avoidance_sweep.py [generations=8] [games=2048] [sims=48]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402,F401
import alphazero_chess_b200 as az  # noqa: E402
from alphazero_chess_b200 import training as tr  # noqa: E402

G, S = 4096, 800


def measure(weights, label, cache_log2=24, max_iters=None, warm_plies=2):
    if max_iters:
        os.environ["AZ_ADV_MAX_ITERS"] = str(max_iters)
    eng = az.Engine(max_games=G, num_simulations=S, seed=42, cache_log2=cache_log2)
    os.environ.pop("AZ_ADV_MAX_ITERS", None)
    eng.load_weights(weights)
    eng.selfplay_begin(G)
    for _ in range(warm_plies):
        eng.selfplay_step(S)
    st0 = eng.selfplay_step(0)
    eng.timer_start()
    st1 = eng.selfplay_step(3 * S)
    ms = eng.timer_stop()
    d = {k: getattr(st1, k) - getattr(st0, k) for k in ("simulations", "evaluations", "cache_hits", "cache_evictions", "terminal_leaves", "sum_leaf_depth")}
    eng.close()
    out = {"label": label, "warm_plies": warm_plies, "cache_log2": cache_log2, "max_iters": max_iters, "sims_per_sec": d["simulations"] / ms * 1e3,
           "evals_per_sec": d["evaluations"] / ms * 1e3, "avoidance": 1 - d["evaluations"] / d["simulations"],
           "hits": d["cache_hits"], "evictions": d["cache_evictions"], "terminal": d["terminal_leaves"],
           "mean_leaf_depth": d["sum_leaf_depth"] / d["simulations"], "us_per_wave": ms * 1e3 / (3 * S)}
    print(json.dumps(out), flush=True)
    return out


def sharpen(weights, k):
    names = az.weight_names()
    w = [a.copy() for a in weights]
    for i, n in enumerate(names):
        if n.startswith("policy_conv_2."):
            w[i] = w[i] * np.float32(k)
    return w


def main():
    gens = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    games = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    sims = int(sys.argv[3]) if len(sys.argv) > 3 else 48
    torch.manual_seed(42)
    dev = torch.device("cuda", 0)
    w0 = az.random_weights(seed=42)
    measure(w0, "random-init", cache_log2=0)
    measure(w0, "random-init")
    eng = az.Engine(max_games=games, num_simulations=sims, seed=7, num_fullmoves=60)
    model = tr.import_weights(tr.AlphaZeroNet(), w0).to(dev)
    opt = tr.make_optimizer(model)
    replay = az.ReplayBuffer(eng, capacity=100_000, max_batch=tr.BATCH_SIZE)
    t0 = time.perf_counter()
    for it in range(gens):
        m = tr.run_generation(eng, replay, model, opt, it, games, min_replay_size=5000, num_steps=60)
        print(json.dumps({"generation": it, "seconds": time.perf_counter() - t0, "positions": m["positions"],
                          "policy_loss": m.get("avg_policy_loss"), "value_loss": m.get("avg_value_loss")}), flush=True)
        if it in (2, gens - 1):
            w = tr.export_weights(model)
            measure(w, f"trained {it + 1} generations")
    w = tr.export_weights(model)
    replay.close()
    eng.close()
    for k in (2.0, 4.0, 8.0):
        measure(sharpen(w, k), f"trained {gens} generations, policy logits x {k:g}")
    measure(sharpen(w, 4.0), "policy logits x 4, max_iters 8", max_iters=8)
    measure(sharpen(w, 4.0), "policy logits x 4, cache off", cache_log2=0)
    # the same after 12 plies of warm-up: the games have left the shared opening, so hits on OTHER games' evaluations are gone
    for k in (1.0, 4.0, 8.0):
        measure(sharpen(w, k) if k != 1.0 else w, f"trained {gens} generations, policy logits x {k:g}, 12 warm-up plies", warm_plies=12)
    measure(w0, "random-init, 12 warm-up plies", warm_plies=12)


if __name__ == "__main__":
    main()
