#!/usr/bin/env python
"""Short self-play run for ncu captures: 4096 games x `sims` simulations, `plies` plies, bf16 network, no extra legs.
    python tools/bench_short.py [sims] [plies] [cache_log2] [warm_plies]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402,F401
import alphazero_chess_b200 as az  # noqa: E402

sims = int(sys.argv[1]) if len(sys.argv) > 1 else 24
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cache = int(sys.argv[3]) if len(sys.argv) > 3 else 0
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 0
with az.Engine(max_games=4096, num_simulations=sims, seed=42, cache_log2=cache) as e:
    e.load_weights(az.random_weights(seed=42))
    e.selfplay_begin(4096)
    if warm:
        st = e.selfplay_step(sims * warm)
        print(f"warm-up: {warm} plies, {st.simulations} simulations")
    e.timer_start()
    st = e.selfplay_step(sims * plies)
    ms = e.timer_stop()
    print(f"{st.simulations} simulations, {st.evaluations} evaluations in {ms:.1f} ms ({st.simulations / ms * 1e3 / 1e6:.2f} M sims/s)")
