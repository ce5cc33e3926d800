"""Training step and generation loop around the engine (training.rs:40-292, 424-440).

Not on the self-play hot path: the network update is ordinary PyTorch (autograd + AdamW on the 10x128 ResNet), the data
comes from the device-resident replay buffer and the new weights go back into the engine through az_load_weights.
Reference semantics restated here (burn 0.18 defaults recalled, not pinned by any reference test):
  loss        = mean_b( -sum_i pi[b,i] * log(p[b,i] + 1e-5) ) + 0.5 * mean_b( (v[b] - z[b])^2 )      training.rs:277-292
  optimizer   = AdamW(beta 0.9/0.999, eps 1e-5, weight decay 1e-4), gradients clipped by value to +-1    training.rs:64-67
  learning rate = triangular cycle 1e-3 <-> 1e-2 over 20 iterations, x0.1 every 1000 iterations            training.rs:424-440
  40 steps of 512 samples per iteration once the buffer holds 20,000 positions                            parameters.rs:11,16-17
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

NUM_RES_BLOCKS, NUM_FILTERS = 10, 128           # parameters.rs:7-8
NUM_TRAIN_STEPS, BATCH_SIZE = 40, 512           # parameters.rs:16-17
MIN_REPLAY_SIZE = 20_000                        # parameters.rs:11
BASE_LEARNING_RATE, MAX_LEARNING_RATE = 1e-3, 1e-2
FULL_CYCLE, HALF_CYCLE, DECAY_INTERVAL = 20, 10, 1000
VALUE_LOSS_WEIGHT, WEIGHT_DECAY = 0.5, 1e-4


class ResidualBlock(nn.Module):                 # agent.rs:11-46
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(NUM_FILTERS, NUM_FILTERS, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(NUM_FILTERS, eps=1e-5, momentum=0.1)
        self.conv2 = nn.Conv2d(NUM_FILTERS, NUM_FILTERS, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(NUM_FILTERS, eps=1e-5, momentum=0.1)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + x)


class AlphaZeroNet(nn.Module):                  # agent.rs:49-144
    def __init__(self):
        super().__init__()
        self.input_conv = nn.Conv2d(19, NUM_FILTERS, 3, padding=1)
        self.input_bn = nn.BatchNorm2d(NUM_FILTERS)
        self.res_blocks = nn.ModuleList([ResidualBlock() for _ in range(NUM_RES_BLOCKS)])
        self.policy_conv_1 = nn.Conv2d(NUM_FILTERS, 32, 1)
        self.policy_bn = nn.BatchNorm2d(32)
        self.policy_conv_2 = nn.Conv2d(32, 64, 1)
        self.value_conv = nn.Conv2d(NUM_FILTERS, 8, 1)
        self.value_bn = nn.BatchNorm2d(8)
        self.value_linear_1 = nn.Linear(512, 64)
        self.value_linear_2 = nn.Linear(64, 1)

    def forward(self, x):
        x = F.relu(self.input_bn(self.input_conv(x)))
        for blk in self.res_blocks:
            x = blk(x)
        p = F.relu(self.policy_bn(self.policy_conv_1(x)))
        policy = torch.softmax(self.policy_conv_2(p).flatten(1), dim=1)
        v = F.relu(self.value_bn(self.value_conv(x))).flatten(1)
        v = F.relu(self.value_linear_1(v))
        value = torch.tanh(self.value_linear_2(v)).squeeze(1)
        return policy, value


def _modules_in_weight_order(model):
    """(kind, module) in the order of az_weight_name: conv/linear -> weight, bias; bn -> gamma, beta, mean, var."""
    out = [("conv", model.input_conv), ("bn", model.input_bn)]
    for blk in model.res_blocks:
        out += [("conv", blk.conv1), ("bn", blk.bn1), ("conv", blk.conv2), ("bn", blk.bn2)]
    out += [("conv", model.policy_conv_1), ("bn", model.policy_bn), ("conv", model.policy_conv_2), ("conv", model.value_conv),
            ("bn", model.value_bn), ("linear", model.value_linear_1), ("linear", model.value_linear_2)]
    return out


def export_weights(model):
    """The 144 f32 arrays az_load_weights expects (burn layout: Linear weight is [d_in, d_out])."""
    arrays = []
    for kind, m in _modules_in_weight_order(model):
        if kind == "bn":
            arrays += [m.weight, m.bias, m.running_mean, m.running_var]
        elif kind == "linear":
            arrays += [m.weight.t(), m.bias]
        else:
            arrays += [m.weight, m.bias]
    return [a.detach().float().cpu().contiguous().numpy().ravel().copy() for a in arrays]


def import_weights(model, arrays):
    it = iter(arrays)
    with torch.no_grad():
        for kind, m in _modules_in_weight_order(model):
            if kind == "bn":
                for t in (m.weight, m.bias, m.running_mean, m.running_var):
                    t.copy_(torch.from_numpy(np.asarray(next(it), np.float32)).view_as(t))
            elif kind == "linear":
                w = torch.from_numpy(np.asarray(next(it), np.float32)).view(m.in_features, m.out_features)
                m.weight.copy_(w.t())
                m.bias.copy_(torch.from_numpy(np.asarray(next(it), np.float32)))
            else:
                m.weight.copy_(torch.from_numpy(np.asarray(next(it), np.float32)).view_as(m.weight))
                m.bias.copy_(torch.from_numpy(np.asarray(next(it), np.float32)))
    return model


def compute_loss(predicted_policy, target_policy, predicted_value, target_value):
    """compute_gradients (training.rs:277-292) without the backward call."""
    difference = predicted_value - target_value
    policy_loss = -(target_policy * (predicted_policy + 1e-5).log()).sum(1).mean()
    value_loss = (difference * difference).mean()
    return policy_loss, value_loss, policy_loss + value_loss * VALUE_LOSS_WEIGHT


def get_cyclical_lr(iteration):
    """training.rs:424-440."""
    decay_multiplier = 10.0 ** (-(iteration // DECAY_INTERVAL))
    base_lr, max_lr = BASE_LEARNING_RATE * decay_multiplier, MAX_LEARNING_RATE * decay_multiplier
    cur = iteration % FULL_CYCLE
    lr_range = max_lr - base_lr
    if cur <= HALF_CYCLE:
        return base_lr + (cur / HALF_CYCLE) * lr_range
    return max_lr - ((cur - HALF_CYCLE) / HALF_CYCLE) * lr_range


def make_optimizer(model):
    return torch.optim.AdamW(model.parameters(), lr=BASE_LEARNING_RATE, betas=(0.9, 0.999), eps=1e-5, weight_decay=WEIGHT_DECAY)


def train_iteration(model, optimizer, replay, iteration, num_steps=NUM_TRAIN_STEPS, batch_size=BATCH_SIZE, seed=0):
    """The inner loop of train() (training.rs:137-200): num_steps batches from the replay buffer, one AdamW step each.
    Returns (avg policy loss, avg value loss)."""
    device = next(model.parameters()).device
    model.train()
    lr = get_cyclical_lr(iteration)
    for g in optimizer.param_groups:
        g["lr"] = lr
    tot_p = tot_v = 0.0
    for step in range(num_steps):
        planes, policy, value = replay.sample(batch_size, seed=(seed * 1_000_003 + iteration) * 65_537 + step)
        if planes.shape[0] == 0:
            break
        x = torch.from_numpy(planes).to(device)
        pi = torch.from_numpy(policy).to(device)
        z = torch.from_numpy(value).to(device)
        p, v = model(x)
        pl, vl, loss = compute_loss(p, pi, v, z)
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_value_(model.parameters(), 1.0)
        optimizer.step()
        tot_p += float(pl.detach())
        tot_v += float(vl.detach())
    n = max(num_steps, 1)
    return tot_p / n, tot_v / n


def run_generation(engine, replay, model, optimizer, iteration, n_games, min_replay_size=MIN_REPLAY_SIZE, waves_per_call=64,
                   num_steps=NUM_TRAIN_STEPS, batch_size=BATCH_SIZE):
    """One iteration of train() (training.rs:70-200) on the engine: self-play until n_games games have finished, their
    steps go from device memory straight into the replay buffer, then the training steps, then the new weights are
    loaded into the engine.  Returns a dict of the metrics the reference logs."""
    engine.load_weights(export_weights(model))
    engine.selfplay_begin(n_games, first_game_id=iteration * (1 << 24))
    new_unique = steps = 0
    while True:
        st = engine.selfplay_step(waves_per_call)
        if st.pending_samples:
            n, nu = replay.add_pending()
            steps += n
            new_unique += nu
        if st.games_finished >= n_games:
            break
    out = {"iteration": iteration, "positions": steps, "new_unique_states": new_unique, "replay_buffer_size": len(replay),
           "simulations": int(st.simulations), "evaluations": int(st.evaluations), "trained": False}
    if len(replay) >= min_replay_size:
        pl, vl = train_iteration(model, optimizer, replay, iteration, num_steps, batch_size)
        out.update(trained=True, avg_policy_loss=pl, avg_value_loss=vl, learning_rate=get_cyclical_lr(iteration))
        engine.load_weights(export_weights(model))
    return out


def run_generation_sharded(engine, replay, model, optimizer, iteration, games_per_rank, dist, device, min_replay_size=MIN_REPLAY_SIZE,
                           waves_per_call=64, num_steps=NUM_TRAIN_STEPS, batch_size=BATCH_SIZE):
    """run_generation over several GPUs (BASELINE config 5): every rank plays its own games (disjoint game ids,
    no data-path collective), the finished games' steps are gathered to rank 0 in rank order and added to ITS replay
    buffer (the single FEN-keyed buffer of memory.rs), rank 0 runs the training steps, and the new weights go back to all
    ranks with one broadcast and an on-device import.  `replay`, `model` and `optimizer` are only used on rank 0."""
    from . import SAMPLE_DTYPE, sharding, weight_sizes

    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    sizes = weight_sizes()
    offs = sharding.weight_offsets(sizes)
    flat = torch.zeros(int(offs[-1]), dtype=torch.float32, device=device)

    def sync_weights():
        if rank == 0:
            flat.copy_(torch.from_numpy(sharding.flatten_weights(export_weights(model))))
        sharding.broadcast_weights(flat, dist, src=0)
        if flat.is_cuda:
            torch.cuda.synchronize(flat.device)
            engine.load_weights_dev([flat.data_ptr() + 4 * int(o) for o in offs[:-1]])
        else:
            engine.load_weights(sharding.split_weights(flat.numpy(), sizes))

    sync_weights()
    engine.selfplay_begin(games_per_rank, first_game_id=sharding.first_game_id(rank) + iteration * (1 << 24))
    steps = new_unique = 0
    done = False
    st = None
    while True:
        mine = np.zeros(0, SAMPLE_DTYPE)
        if not done:
            st = engine.selfplay_step(waves_per_call)
            if st.pending_samples:
                mine = engine.selfplay_drain()
            done = st.games_finished >= games_per_rank
        got = sharding.gather_samples(mine, dist, device if flat.is_cuda else None)
        if rank == 0 and len(got):
            steps += len(got)
            new_unique += replay.add(got)
        if sharding.all_done(done, dist, device if flat.is_cuda else None):
            break
    sums, _ = sharding.reduce_metrics([float(st.simulations), float(st.evaluations)], [0.0], dist)
    out = {"iteration": iteration, "n_ranks": world, "positions": steps, "new_unique_states": new_unique,
           "replay_buffer_size": len(replay) if rank == 0 else 0, "simulations": int(sums[0]), "evaluations": int(sums[1]), "trained": False}
    train = all_flag = False
    if rank == 0:
        train = len(replay) >= min_replay_size
    t = torch.tensor([1 if train else 0], dtype=torch.int32, device=device if flat.is_cuda else None)
    if world > 1:
        dist.broadcast(t, src=0)
    all_flag = bool(t.item())
    if all_flag:
        if rank == 0:
            pl, vl = train_iteration(model, optimizer, replay, iteration, num_steps, batch_size)
            out.update(trained=True, avg_policy_loss=pl, avg_value_loss=vl, learning_rate=get_cyclical_lr(iteration))
        sync_weights()
    return out
