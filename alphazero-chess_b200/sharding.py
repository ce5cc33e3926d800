"""Multi-GPU plumbing: games are independent, so every rank owns its games outright (no data-path collective).
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used for exactly two things: the
per-generation weight broadcast and the reduction of counters / timings for reporting."""
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


def all_done(done, dist=None, device=None):
    """True once every rank reports done (one small all-reduce)."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bool(done)
    t = torch.tensor([1 if done else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())
