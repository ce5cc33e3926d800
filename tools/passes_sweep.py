#!/usr/bin/env python
"""Sweep of (network-free simulations per k_advance pass, passes per wave) at high cache hit rates: trains briefly, sharpens
the policy head, then measures BASELINE configs[2] self-play with the cache on for each setting.  One JSON line each.
    python tools/passes_sweep.py [generations=6]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402,F401
import alphazero_chess_b200 as az  # noqa: E402
from alphazero_chess_b200 import training as tr  # noqa: E402

G, S = 4096, 800


def measure(weights, label, max_iters, passes):
    os.environ["AZ_ADV_MAX_ITERS"] = str(max_iters)
    if passes:
        os.environ["AZ_ADV_PASSES"] = str(passes)
    else:
        os.environ.pop("AZ_ADV_PASSES", None)
    eng = az.Engine(max_games=G, num_simulations=S, seed=42, cache_log2=24)
    eng.load_weights(weights)
    eng.selfplay_begin(G)
    for _ in range(2):
        eng.selfplay_step(S)
    st0 = eng.selfplay_step(0)
    eng.timer_start()
    for _ in range(3):
        st1 = eng.selfplay_step(S)
    ms = eng.timer_stop()
    d = {k: getattr(st1, k) - getattr(st0, k) for k in ("simulations", "evaluations")}
    eng.close()
    print(json.dumps({"label": label, "max_iters": max_iters, "passes": passes or "adaptive", "sims_per_sec": d["simulations"] / ms * 1e3,
                      "evals_per_sec": d["evaluations"] / ms * 1e3, "avoidance": 1 - d["evaluations"] / d["simulations"],
                      "us_per_wave": ms * 1e3 / (3 * S)}), flush=True)


def main():
    gens = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    torch.manual_seed(42)
    dev = torch.device("cuda", 0)
    eng = az.Engine(max_games=2048, num_simulations=48, seed=7, num_fullmoves=60)
    model = tr.import_weights(tr.AlphaZeroNet(), az.random_weights(seed=42)).to(dev)
    opt = tr.make_optimizer(model)
    replay = az.ReplayBuffer(eng, capacity=100_000, max_batch=tr.BATCH_SIZE)
    for it in range(gens):
        tr.run_generation(eng, replay, model, opt, it, 2048, min_replay_size=5000, num_steps=60)
    w = tr.export_weights(model)
    replay.close()
    eng.close()
    names = az.weight_names()
    for k in (1.0, 4.0, 8.0):
        ws = [a * np.float32(k) if n.startswith("policy_conv_2.") else a for a, n in zip(w, names)]
        for max_iters, passes in ((4, 1), (4, 2), (4, 3), (2, 2), (2, 3), (3, 3), (4, 0)):
            measure(ws, f"trained {gens} generations, policy logits x {k:g}", max_iters, passes)


if __name__ == "__main__":
    main()
