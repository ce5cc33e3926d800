"""Multi-GPU plumbing: games are independent, so every rank owns its games outright (no data-path collective).
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used at generation boundaries only: the
weight broadcast, the gather of finished games' samples into the replay buffer(s) -- device memory to device memory on
NCCL -- the gradient all-reduce of data-parallel training, and the reduction of counters / timings for reporting."""
import numpy as np

GAME_ID_STRIDE = 1 << 40  # rank r owns game ids [r * 2^40, (r+1) * 2^40): disjoint noise / sampling streams


def first_game_id(rank):
    return int(rank) * GAME_ID_STRIDE


def weight_offsets(sizes):
    """Element offsets of the 144 arrays inside the flat broadcast buffer (last entry = total)."""
    return np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64))]).astype(np.int64)


def flatten_weights(arrays):
    return np.concatenate([np.asarray(a, np.float32).ravel() for a in arrays])


def split_weights(flat, sizes):
    offs = weight_offsets(sizes)
    return [flat[int(offs[i]): int(offs[i + 1])] for i in range(len(sizes))]


def broadcast_weights(flat_tensor, dist=None, src=0):
    """One collective per generation: rank `src` holds the new weights, everybody else receives them in place."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_tensor, src=src)
    return flat_tensor


def reduce_metrics(sums, maxes, dist=None):
    """Whole-job aggregates: work counters are summed over ranks, times are the max over ranks."""
    import torch

    s = torch.as_tensor(sums, dtype=torch.float64).clone()
    m = torch.as_tensor(maxes, dtype=torch.float64).clone()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dev = None
        if dist.get_backend() == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            s, m = s.to(dev), m.to(dev)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        s, m = s.cpu(), m.cpu()
    return s.numpy(), m.numpy()


def gather_samples(samples, dist=None, device=None, dst=0):
    """Replay-buffer gather: every rank hands in the az_sample records of its finished games; rank `dst`
    receives them concatenated in rank order (the order ReplayBuffer::add then sees), the others get an empty array.
    One size exchange plus one padded gather per call; `device` is the CUDA device for the NCCL backend (None = CPU/gloo)."""
    import torch

    samples = np.ascontiguousarray(samples)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return samples
    world, rank = dist.get_world_size(), dist.get_rank()
    item = samples.dtype.itemsize
    n = torch.tensor([samples.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    counts = [int(t.item()) for t in sizes]
    width = max(counts)
    if width == 0:
        return samples[:0]
    buf = torch.zeros(width * item, dtype=torch.uint8, device=device)
    if samples.shape[0]:
        buf[: samples.shape[0] * item] = torch.from_numpy(samples.view(np.uint8).reshape(-1)).to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst)
    if rank != dst:
        return samples[:0]
    out = [parts[r][: counts[r] * item].cpu().numpy().view(samples.dtype) for r in range(world) if counts[r]]
    return np.concatenate(out)


def allgather_samples(samples, dist=None):
    """Host-array version of the sample exchange (gloo / CPU tests): every rank receives all ranks' records in rank order."""
    import torch

    samples = np.ascontiguousarray(samples)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return samples
    world = dist.get_world_size()
    item = samples.dtype.itemsize
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([samples.shape[0]], dtype=torch.int64))
    counts = [int(t.item()) for t in sizes]
    width = max(counts)
    if width == 0:
        return samples[:0]
    buf = torch.zeros(width * item, dtype=torch.uint8)
    if samples.shape[0]:
        buf[: samples.shape[0] * item] = torch.from_numpy(samples.view(np.uint8).reshape(-1))
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = [parts[r][: counts[r] * item].numpy().view(samples.dtype) for r in range(world) if counts[r]]
    return np.concatenate(out)


class DeviceSampleExchange:
    """Replay-buffer gather without a host hop (north_star: NCCL "to gather replay-buffer samples"; training.rs:86-105):
    az_selfplay_drain_dev copies the finished games' az_sample records into a CUDA send buffer, one NCCL all-gather moves
    every rank's records into every rank's receive buffer, and az_replay_add_dev applies them to the local replay buffer in
    rank order -- so all ranks hold the same FEN-keyed buffer (memory.rs:41-76 sees the same sequence of steps everywhere)
    and data-parallel training can sample identical batches.  Only the per-rank record COUNTS travel through the host."""

    def __init__(self, engine, dist, device, max_samples):
        import torch

        from . import SAMPLE_DTYPE

        self.engine, self.dist, self.device = engine, dist, device
        self.item = SAMPLE_DTYPE.itemsize
        self.max_samples = int(max_samples)
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.send = torch.empty(self.max_samples * self.item, dtype=torch.uint8, device=device)
        self.recv = torch.empty(self.world * self.max_samples * self.item, dtype=torch.uint8, device=device) if self.world > 1 else None
        self.counts = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.bytes_moved = 0

    def exchange(self):
        """Returns [(device pointer, n records)] in rank order (one entry when there is a single rank)."""
        import torch

        n = self.engine.selfplay_drain_dev(self.send.data_ptr(), self.max_samples)
        if self.world == 1:
            return [(self.send.data_ptr(), n)]
        mine = torch.tensor([n], dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(self.counts, mine)
        counts = [int(c) for c in self.counts.cpu()]
        width = max(counts)
        if width == 0:
            return [(self.recv.data_ptr(), 0)]
        nbytes = width * self.item
        self.dist.all_gather_into_tensor(self.recv[: self.world * nbytes], self.send[:nbytes])
        torch.cuda.synchronize(self.device)
        self.bytes_moved += self.world * nbytes
        return [(self.recv.data_ptr() + r * nbytes, counts[r]) for r in range(self.world)]


def allreduce_gradients(params, dist, flat=None):
    """Data-parallel training (training.rs:137-200 over several GPUs): one flat all-reduce (sum) of all gradients, 12.1 MB
    for the 10x128 network.  Every rank has already scaled its loss by its share of the global batch, so the sum IS the
    gradient of the reference's full-batch mean loss."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    grads = [p.grad for p in params if p.grad is not None]
    if flat is None or flat.numel() != sum(g.numel() for g in grads):
        flat = torch.empty(sum(g.numel() for g in grads), dtype=grads[0].dtype, device=grads[0].device)
    torch.cat([g.reshape(-1) for g in grads], out=flat)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        g.copy_(flat[off: off + g.numel()].view_as(g))
        off += g.numel()
    return flat


def all_done(done, dist=None, device=None):
    """True once every rank reports done (one small all-reduce)."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bool(done)
    t = torch.tensor([1 if done else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())
