#!/usr/bin/env python
"""Generates tests/golden/movegen.json with the ORACLE: the ORDERED legal-move lists (wire moves and policy indices) and the
input planes digest of the special positions of tests/helpers.py.  The order of `legal_moves()` is restated from memory of
shakmaty's generator ("parity unpinned"); the fixture freezes that restatement so that oracle and
engine cannot change it together unnoticed.

    python tools/make_golden_movegen.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import SPECIAL_FENS, orc  # noqa: E402


def main():
    out = []
    for fen in SPECIAL_FENS:
        p = orc.from_fen(fen)
        mv, idx = orc.legal_moves(p)
        out.append({"fen": fen, "moves": [int(m) for m in mv], "index": [int(i) for i in idx], "outcome": int(orc.outcome(p)),
                    "planes_sha256": hashlib.sha256(orc.to_tensor(p).tobytes()).hexdigest()})
    with open(os.path.join(ROOT, "tests", "golden", "movegen.json"), "w") as f:
        json.dump({"_how": "python tools/make_golden_movegen.py (oracle); move = from | to << 6 | promo << 12 | special << 15",
                   "positions": out}, f, indent=1)
    print("wrote", len(out), "positions")


if __name__ == "__main__":
    main()
