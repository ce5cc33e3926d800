"""Shared test helpers (the oracle is test infrastructure; see oracle/oracle.hpp)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
SPECIAL_FENS = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    KIWIPETE,
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
    "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10",
    "8/8/8/8/k2Pp2Q/8/8/4K3 b - d3 0 1",      # en passant would expose the king along the rank
    "8/8/8/2k5/3Pp3/8/8/4K3 b - d3 0 1",      # en passant capture of a checking pawn
    "4k3/8/8/8/8/8/8/4K2R w K - 0 1",
    "r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1",
    "r3k2r/8/8/8/8/8/6n1/R3K2R w KQkq - 0 1",  # knight attacks f1/h1 region
    "4k3/8/8/8/8/8/4q3/4K3 w - - 0 1",        # in check by an adjacent queen
    "7k/5Q2/6K1/8/8/8/8/8 b - - 0 1",          # stalemate
    "7k/6Q1/6K1/8/8/8/8/8 b - - 0 1",          # checkmate
    "8/8/8/8/8/2k5/8/K1n5 w - - 0 1",          # K+N vs K
    "8/8/8/8/8/2k5/8/KB6 w - - 0 1",
    "1n5k/P7/8/8/8/8/8/K7 w - - 0 1",          # promotions incl. capture
    "k7/8/8/8/8/8/7p/K5N1 b - - 0 1",
    "rnbqkbnr/ppp1pppp/8/3pP3/8/8/PPPP1PPP/RNBQKBNR w KQkq d6 0 3",
    "rnbqkbnr/pppp1ppp/8/8/4pP2/8/PPPPP1PP/RNBQKBNR b KQkq f3 99 150",
]


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def random_playouts(n_positions, seed=42, max_plies=80, roots=None):
    """Positions reached by seeded random playouts, with the history of each."""
    rng = np.random.default_rng(seed)
    roots = roots or [SPECIAL_FENS[0], KIWIPETE]
    positions, histories = [], []
    while len(positions) < n_positions:
        pos = orc.from_fen(roots[int(rng.integers(len(roots)))])
        hist = [pos.copy()]
        plies = int(rng.integers(0, max_plies + 1))
        for _ in range(plies):
            mv, _ix = orc.legal_moves(pos)
            if len(mv) == 0 or orc.outcome(pos) != 0:
                break
            pos = orc.play_encoded(pos, mv[int(rng.integers(len(mv)))])
            hist.append(pos.copy())
        positions.append(pos.copy())
        histories.append(np.array(hist, orc.POSITION_DTYPE))
    return np.array(positions, orc.POSITION_DTYPE), histories


def uci_to_wire(pos, uci):
    """UCI (castling as king-takes-rook) -> wire move, matched against the oracle's legal list."""
    f = (ord(uci[0]) - 97) + 8 * (int(uci[1]) - 1)
    t = (ord(uci[2]) - 97) + 8 * (int(uci[3]) - 1)
    promo = {"n": 1, "b": 2, "r": 3, "q": 4}.get(uci[4:5], 0)
    mv, _ = orc.legal_moves(pos)
    for m in mv:
        if (m & 63) == f and ((m >> 6) & 63) == t and ((m >> 12) & 7) == promo:
            return int(m)
    raise AssertionError(f"{uci} is not legal")


def torch_reference_forward(arrays, planes):
    """Plain PyTorch fp32 restatement of AlphaZero::forward (agent.rs:112-144) from the 144 burn-layout arrays."""
    import torch
    import torch.nn.functional as F

    a = [torch.from_numpy(np.asarray(x, np.float32)) for x in arrays]
    x = torch.from_numpy(np.asarray(planes, np.float32)).reshape(-1, 19, 8, 8)

    def bn(x, i):
        return F.batch_norm(x, a[i + 2], a[i + 3], a[i], a[i + 1], training=False, eps=1e-5)  # arrays: gamma, beta, mean, var

    with torch.no_grad():
        x = torch.relu(bn(F.conv2d(x, a[0].view(128, 19, 3, 3), a[1], padding=1), 2))
        for b in range(10):
            o = 6 + b * 12
            r = x
            y = torch.relu(bn(F.conv2d(x, a[o].view(128, 128, 3, 3), a[o + 1], padding=1), o + 2))
            y = bn(F.conv2d(y, a[o + 6].view(128, 128, 3, 3), a[o + 7], padding=1), o + 8)
            x = torch.relu(y + r)
        h = 6 + 120
        p = torch.relu(bn(F.conv2d(x, a[h].view(32, 128, 1, 1), a[h + 1]), h + 2))
        logits = F.conv2d(p, a[h + 6].view(64, 32, 1, 1), a[h + 7]).reshape(x.shape[0], -1)
        policy = torch.softmax(logits, 1)
        v = torch.relu(bn(F.conv2d(x, a[h + 8].view(8, 128, 1, 1), a[h + 9]), h + 10)).reshape(x.shape[0], -1)
        v = torch.relu(v @ a[h + 14].view(512, 64) + a[h + 15])
        v = torch.tanh(v @ a[h + 16].view(64, 1) + a[h + 17]).squeeze(1)
    return policy.numpy(), v.numpy()
