"""training.rs pieces around the engine: learning-rate schedule, loss, weight import/export."""
import numpy as np
import torch

import _pkg  # noqa: F401  (spawned workers re-import this module without conftest)
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


def test_training_step_matches_numpy_oracle():
    """The PyTorch training step (loss, value clipping, AdamW, cyclical LR) against oracle/train_step.py (float64 numpy
    restatement of training.rs:64-67,277-292,424-440) on the real 10x128 network: three consecutive steps."""
    from oracle import train_step as ots

    torch.manual_seed(0)
    model = tr.AlphaZeroNet()
    opt = tr.make_optimizer(model)
    params = [p for p in model.parameters()]
    ora = ots.AdamW([tuple(p.shape) for p in params])
    w = [p.detach().double().numpy().copy() for p in params]
    rng = np.random.default_rng(5)
    for step, iteration in enumerate((0, 7, 1013)):
        assert abs(tr.get_cyclical_lr(iteration) - ots.cyclical_lr(iteration)) < 1e-15
        x = torch.from_numpy(rng.uniform(0, 1, (16, 19, 8, 8)).astype(np.float32))
        pi = torch.from_numpy(rng.dirichlet(np.ones(4096) * 0.02, 16).astype(np.float32))
        z = torch.from_numpy(rng.uniform(-1, 1, 16).astype(np.float32))
        model.train()
        p, v = model(x)
        pl, vl, loss = tr.compute_loss(p, pi, v, z)
        opl, ovl, oloss = ots.loss(p.detach().numpy(), pi.numpy(), v.detach().numpy(), z.numpy())
        assert abs(float(pl.detach()) - opl) < 1e-4 and abs(float(vl.detach()) - ovl) < 1e-6 and abs(float(loss.detach()) - oloss) < 1e-4
        # the gradients autograd starts from
        p.retain_grad(), v.retain_grad()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        gp, gv = ots.loss_gradients(p.detach().numpy(), pi.numpy(), v.detach().numpy(), z.numpy())
        assert np.allclose(p.grad.numpy(), gp, rtol=1e-4, atol=1e-7) and np.allclose(v.grad.numpy(), gv, rtol=1e-4, atol=1e-8)
        grads = [q.grad.detach().double().numpy().copy() for q in params]
        for g in opt.param_groups:
            g["lr"] = tr.get_cyclical_lr(iteration)
        torch.nn.utils.clip_grad_value_(params, 1.0)
        opt.step()
        w = ora.step(w, grads, ots.cyclical_lr(iteration))
        worst = max(float(np.abs(q.detach().double().numpy() - wi).max()) for q, wi in zip(params, w))
        assert worst < 2e-6, (step, worst)
        w = [q.detach().double().numpy().copy() for q in params]   # re-base so f32 rounding does not accumulate over steps


class _TinyNet(torch.nn.Module):
    """BatchNorm-free stand-in with the network's interface, so that full-batch and data-parallel updates agree to rounding."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(19, 8, 3, padding=1)
        self.pol = torch.nn.Linear(8 * 64, 4096)
        self.val = torch.nn.Linear(8 * 64, 1)

    def forward(self, x):
        h = torch.relu(self.conv(x)).flatten(1)
        return torch.softmax(self.pol(h), 1), torch.tanh(self.val(h)).squeeze(1)


class _FakeReplay:
    def __init__(self, n=64):
        rng = np.random.default_rng(11)
        self.planes = rng.uniform(0, 1, (n, 19, 8, 8)).astype(np.float32)
        self.policy = rng.dirichlet(np.ones(4096) * 0.05, n).astype(np.float32)
        self.value = rng.uniform(-1, 1, n).astype(np.float32)

    def sample(self, batch_size, seed=0):
        idx = np.random.default_rng(seed).permutation(len(self.value))[:batch_size]
        return self.planes[idx], self.policy[idx], self.value[idx]


def _dp_worker(rank, world, port, out_dir):
    import os

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    model = _TinyNet()
    opt = tr.make_optimizer(model)
    pl, vl = tr.train_iteration(model, opt, _FakeReplay(), iteration=4, num_steps=3, batch_size=32, dist=dist)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    np.save(os.path.join(out_dir, f"dp{rank}.npy"), np.concatenate([flat, [pl, vl]]))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_training_equals_full_batch(tmp_path):
    """train_iteration over two gloo ranks (rows rank::2 each, one gradient all-reduce per step) applies the same updates as
    one process on the full batch: training.rs:137-200 made data parallel without changing its arithmetic."""
    import os

    import torch.multiprocessing as mp

    torch.manual_seed(3)
    model = _TinyNet()
    opt = tr.make_optimizer(model)
    pl, vl = tr.train_iteration(model, opt, _FakeReplay(), iteration=4, num_steps=3, batch_size=32)
    want = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    world, port = 2, 33500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "dp0.npy"), np.load(tmp_path / "dp1.npy")
    assert np.array_equal(r0, r1)                                  # replicas stay identical
    assert np.abs(r0[:-2] - want).max() < 1e-5                     # and equal the single-process full-batch result
    assert abs(r0[-2] - pl) < 1e-4 and abs(r0[-1] - vl) < 1e-5
