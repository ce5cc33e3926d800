"""training.rs pieces around the engine: learning-rate schedule, loss, weight import/export."""
import numpy as np
import torch

import alphazero_chess_b200 as az
from alphazero_chess_b200 import training as tr
from helpers import orc, random_playouts, torch_reference_forward


def test_cyclical_lr_matches_formula():
    # training.rs:424-440: 1e-3 -> 1e-2 over 10 iterations, back over the next 10, x0.1 every 1000
    assert abs(tr.get_cyclical_lr(0) - 1e-3) < 1e-12
    assert abs(tr.get_cyclical_lr(10) - 1e-2) < 1e-12
    assert abs(tr.get_cyclical_lr(5) - 5.5e-3) < 1e-12
    assert abs(tr.get_cyclical_lr(15) - 5.5e-3) < 1e-12
    assert abs(tr.get_cyclical_lr(20) - 1e-3) < 1e-12
    assert abs(tr.get_cyclical_lr(1010) - 1e-3) < 1e-12
    assert abs(tr.get_cyclical_lr(2005) - 5.5e-5) < 1e-14


def test_loss_matches_manual():
    rng = np.random.default_rng(0)
    p = rng.dirichlet(np.ones(4096), 8).astype(np.float32)
    pi = rng.dirichlet(np.ones(4096) * 0.01, 8).astype(np.float32)
    v = rng.uniform(-1, 1, 8).astype(np.float32)
    z = rng.uniform(-1, 1, 8).astype(np.float32)
    pl, vl, loss = tr.compute_loss(torch.from_numpy(p), torch.from_numpy(pi), torch.from_numpy(v), torch.from_numpy(z))
    want_pl = float(np.mean(-(pi.astype(np.float64) * np.log(p.astype(np.float64) + 1e-5)).sum(1)))
    want_vl = float(np.mean((v.astype(np.float64) - z) ** 2))
    assert abs(float(pl) - want_pl) < 1e-4 and abs(float(vl) - want_vl) < 1e-6
    assert abs(float(loss) - (want_pl + 0.5 * want_vl)) < 1e-4


def test_weight_round_trip_and_forward_parity():
    w = az.random_weights(seed=9, randomize_bn=True)
    model = tr.import_weights(tr.AlphaZeroNet(), w).eval()
    back = tr.export_weights(model)
    assert all(np.array_equal(a, b) for a, b in zip(w, back))
    positions, _ = random_playouts(6, seed=4, max_plies=40)
    planes = np.stack([orc.to_tensor(p) for p in positions])
    with torch.no_grad():
        p, v = model(torch.from_numpy(planes))
    rp, rv = torch_reference_forward(w, planes)
    assert np.abs(p.numpy() - rp).max() < 1e-6 and np.abs(v.numpy() - rv).max() < 1e-6
    # the oracle's C++ network, built from the same arrays, agrees too (pins the burn-layout mapping on both sides)
    op, ov = orc.Net(w).forward_planes(planes)
    assert np.abs(op - rp).max() < 1e-5 and np.abs(ov - rv).max() < 1e-5
