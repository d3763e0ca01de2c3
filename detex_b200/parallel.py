"""parallel.py -- multi-GPU sharding of the hot path (one process per GPU).

The path shards by independent units: detection / FAS by continuous-data
chunk, CCX by row block.  There is NO data-path collective; NCCL (or gloo on CPU for the
tests) is used once at the end to gather the variable-length trigger lists, sum the
per-subspace histograms / FAS sufficient statistics and gather the CCX row blocks.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(nunits, rank, world):
    """Contiguous, balanced [lo, hi) of `nunits` independent units for `rank`."""
    base, rem = divmod(nunits, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ccx_row_blocks(N, world):
    """Row ranges [b0, b1) of the upper-triangular pair matrix with ~equal pair counts.
    Row b owns N-1-b pairs."""
    total = N * (N - 1) // 2
    bounds = [0]
    acc = 0
    b = 0
    for r in range(1, world):
        target = total * r / float(world)
        while b < N - 1 and acc + (N - 1 - b) <= target:
            acc += N - 1 - b
            b += 1
        bounds.append(b)
    bounds.append(N - 1)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_sum(arr):
    """Sum a NumPy array (int64 histogram / float64 statistics) over ranks."""
    if _world() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_dev())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(arr):
    if _world() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def gather_records(rec):
    """All-gather a 1-D structured NumPy array of variable length (trigger / candidate
    records): counts first, then the padded byte payload.  Returns the concatenation in
    rank order on every rank."""
    world = _world()
    if world == 1:
        return rec
    dev = _dev()
    n = torch.tensor([len(rec)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    item = rec.dtype.itemsize
    buf = np.zeros(mx * item, dtype=np.uint8)
    buf[:len(rec) * item] = np.frombuffer(np.ascontiguousarray(rec).tobytes(), dtype=np.uint8)
    t = torch.from_numpy(buf).to(dev)
    outs = [torch.zeros(mx * item, dtype=torch.uint8, device=dev) for _ in range(world)]
    if mx > 0:
        dist.all_gather(outs, t)
    parts = [np.frombuffer(o.cpu().numpy().tobytes()[:c * item], dtype=rec.dtype) for o, c in zip(outs, counts)]
    return np.concatenate(parts) if parts else rec


def gather_row_blocks(block, blocks, N):
    """All-gather CCX row blocks (rows_i, N) into the full (N-1, N) matrix on every rank."""
    world = _world()
    if world == 1:
        return block
    dev = _dev()
    mx = max(b1 - b0 for b0, b1 in blocks)
    pad = np.zeros((mx, N), dtype=block.dtype)
    pad[:block.shape[0]] = block
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:b1 - b0] for o, (b0, b1) in zip(outs, blocks)], axis=0)
