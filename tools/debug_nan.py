#!/usr/bin/env python
"""Where do non-finite training losses come from?  Replays the bench's peaked_weights() recipe and checks every stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa
import numpy as np, torch
import alphazero_chess_b200 as az
from alphazero_chess_b200 import training as tr

torch.manual_seed(42)
dev = torch.device("cuda", 0)
eng = az.Engine(device=0, max_games=2048, num_simulations=48, seed=7, num_fullmoves=60)
model = tr.import_weights(tr.AlphaZeroNet(), az.random_weights(seed=42)).to(dev)
opt = tr.make_optimizer(model)
replay = az.ReplayBuffer(eng, capacity=100_000, max_batch=tr.BATCH_SIZE)
m = tr.run_generation(eng, replay, model, opt, 0, 2048, min_replay_size=10**9, num_steps=0)
print("generation 0:", {k: m[k] for k in ("positions", "new_unique_states", "replay_buffer_size", "evaluations")})
bad = 0
for s in range(20):
    planes, policy, value = replay.sample_torch(tr.BATCH_SIZE, seed=s)
    fin = torch.isfinite(planes).all().item(), torch.isfinite(policy).all().item(), torch.isfinite(value).all().item()
    ps = policy.sum(1)
    if not all(fin) or (ps - 1).abs().max().item() > 1e-3:
        bad += 1
        print("batch", s, "finite planes/policy/value", fin, "policy row sums", ps.min().item(), ps.max().item(), "value range", value.min().item(), value.max().item())
print("bad batches:", bad)
model.train()
for step in range(60):
    planes, policy, value = replay.sample_torch(tr.BATCH_SIZE, seed=1000 + step)
    p, v = model(planes)
    pl = -(policy * (p + 1e-5).log()).sum() / planes.shape[0]
    vl = ((v - value) ** 2).sum() / planes.shape[0]
    loss = pl + vl * tr.VALUE_LOSS_WEIGHT
    opt.zero_grad(set_to_none=True)
    loss.backward()
    gn = torch.sqrt(sum((q.grad.double() ** 2).sum() for q in model.parameters() if q.grad is not None)).item()
    if step < 3 or not np.isfinite(loss.item()) or step % 20 == 0:
        print(f"step {step}: pl {pl.item():.4f} vl {vl.item():.5f} |grad| {gn:.3e} p finite {torch.isfinite(p).all().item()} v finite {torch.isfinite(v).all().item()}")
    if not np.isfinite(loss.item()):
        break
    torch.nn.utils.clip_grad_value_(model.parameters(), 1.0) if hasattr(tr, "GRAD_CLIP") else None
    opt.step()
