"""The C++ host-side mirror (include/az_b200.hpp) driven like the reference's callers; results checked against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import ROOT, orc

pytestmark = pytest.mark.gpu


def test_cpp_mirror(tmp_path):
    exe = str(tmp_path / "test_mirror")
    lib_dir = os.path.dirname(az.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"),
                           "-o", exe, "-L", lib_dir, "-laz_b200", f"-Wl,-rpath,{lib_dir}"])
    stub_seed = 5
    out = subprocess.run([exe, str(stub_seed)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    res = {d["test"]: d for d in (json.loads(l) for l in out.stdout.splitlines() if l.startswith("{"))}
    assert res["repetition"] == {"test": "repetition", "plies": 8, "result": az.DRAW, "counted": 9}
    assert res["illegal"]["result"] == az.ILLEGAL and res["illegal"]["none"] == 1 and res["illegal"]["e2e4"] == 588
    assert abs(res["planes"]["sum"] - (32 + 4 * 64 + 64 * (1 / 200))) < 1e-3  # f32 accumulation in the C++ program
    prm = orc.make_params(num_simulations=48)
    ev = orc.make_evaluator("stub", stub_seed=stub_seed)
    v, _, d, _ = orc.search(orc.startpos(), prm, ev)
    got = np.zeros(4096, np.float32)
    for i, c in res["search"]["visits"]:
        got[i] = c
    assert np.array_equal(got, v) and res["search"]["depth"] == d
    # traverse_new + second search: the oracle searches the position after the move with the history so far
    action = res["search2"]["action"]
    assert action == int(np.flatnonzero(v == v.max()).max())
    p1, r = orc.play_move(orc.startpos(), action, np.array([orc.startpos()], orc.POSITION_DTYPE))
    assert r == 0
    v2, _, d2, _ = orc.search(p1, prm, ev, history=np.array([orc.startpos(), p1], orc.POSITION_DTYPE))
    got2 = np.zeros(4096, np.float32)
    for i, c in res["search2"]["visits"]:
        got2[i] = c
    assert np.array_equal(got2, v2) and res["search2"]["depth"] == d2
    # run_all_episodes: the four games equal the oracle's episodes 0..3
    steps, vsum, dsum = 0, 0.0, 0
    for g in range(4):
        ep = orc.selfplay_episode(prm, ev, game_id=g, max_steps=512, want_visits=False)
        steps += ep["stats"].n_steps
        vsum += float(ep["final_value"].astype(np.float64).sum())
        dsum += int(ep["depth"].sum())
    assert res["episodes"]["steps"] == steps and res["episodes"]["depth_sum"] == dsum
    assert abs(res["episodes"]["value_sum"] - vsum) < 1e-4
    # memory.rs through the mirror: every step was added, duplicates merged, a full batch of probability rows sampled
    rp = res["replay"]
    assert rp["added"] > 8 * 20 and 0 < rp["unique"] <= rp["added"] and rp["len"] == min(rp["unique"], 1000)
    assert rp["batch"] == min(64, rp["len"]) and abs(rp["policy_sum"] - rp["batch"]) < 1e-2
    # chess.rs get_best_move: Ra1-a8 mates; a mated side has no move
    assert res["minimax"] == {"test": "minimax", "from": 0, "to": 56, "none": 1}
